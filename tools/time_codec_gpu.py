#!/usr/bin/env python
"""Device time of the GPU half of compress (``encode_symbols``) per trunk precision, with a per-layer
breakdown of the convolution launches (CUDA events around every launch).

    python tools/time_codec_gpu.py [--tiles 8] [--h 704] [--w 512] [--modes bf16,fp32x2,fp32x3] [--layers]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hyres_b200  # noqa: E402
from hyres_b200 import ops, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", type=int, default=8)
    ap.add_argument("--h", type=int, default=704)
    ap.add_argument("--w", type=int, default=512)
    ap.add_argument("--modes", default="bf16,fp32x2,fp32x3")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--layers", action="store_true")
    a = ap.parse_args()
    torch.manual_seed(1926)
    net = hyres_b200.ResidualJPEGCompression(jpeg_quality=1)
    net.update(force=True)
    net = net.cuda().eval()
    codec = net.residual_model
    x = synthetic.synthetic_image(a.tiles, a.h, a.w, seed=7).cuda()
    jd = (x * 0.9 + 0.05).contiguous()
    px = a.tiles * a.h * a.w
    for mode in a.modes.split(","):
        codec.codec_precision = mode
        with torch.no_grad():
            for _ in range(2):
                codec.encode_symbols(x, _jpeg=jd)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.reps):
                codec.encode_symbols(x, _jpeg=jd)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.reps
            ops.ConvLayer.profile_begin()
            codec.encode_symbols(x, _jpeg=jd)
            conv_ms, conv_n = ops.ConvLayer.profile_end()
        out = dict(mode=mode, nacc=os.environ.get("HYRES_SPLIT_NACC", "default"), tiles=a.tiles, h=a.h, w=a.w,
                   encode_symbols_ms=ms, mpixel_per_s=px / ms / 1e3, conv_ms=conv_ms, conv_launches=conv_n)
        print(json.dumps(out), flush=True)
        if a.layers:
            agg = {}
            for r in ops.ConvLayer.last_profile:
                key = (r["kind"], r["cin"], r["cout"], r["k"], r["stride"], r["H"], r["W"])
                t = agg.setdefault(key, [0, 0.0])
                t[0] += 1
                t[1] += r["ms"]
            for key, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
                print("   ", key, n, "launches", round(t, 3), "ms", flush=True)


if __name__ == "__main__":
    main()
