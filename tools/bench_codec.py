#!/usr/bin/env python
"""compress / decompress timing of the drop-in API (BASELINE.json configs[2]: one 2048x1408 CLIC-shape image as
eight 704x512 tiles, tier T-A of SURVEY section 8e) -- GPU front-end + host rANS, byte strings out and back.

    python tools/bench_codec.py [--tiles 8] [--reps 3] [--oracle]
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hyres_b200  # noqa: E402
from hyres_b200 import synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", type=int, default=8)
    ap.add_argument("--h", type=int, default=704)
    ap.add_argument("--w", type=int, default=512)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    torch.manual_seed(1926)
    net = hyres_b200.ResidualJPEGCompression(jpeg_quality=1)
    net.update(force=True)
    net = net.cuda().eval()
    x = synthetic.synthetic_image(a.tiles, a.h, a.w, seed=7)
    bufs = net.jpeg.compress(x)
    xd = x.cuda()
    px = a.tiles * a.h * a.w
    res = {}
    with torch.no_grad():
        for _ in range(2):
            c = net.compress(xd, jpeg_buffers=bufs)
            d = net.decompress(c)
        torch.cuda.synchronize()
        t_enc = t_dec = 0.0
        for _ in range(a.reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            c = net.compress(xd, jpeg_buffers=bufs)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            d = net.decompress(c)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            t_enc += t1 - t0
            t_dec += t2 - t1
        t_enc /= a.reps
        t_dec /= a.reps
        # GPU-only part of the encoder (symbols of both passes, no host coding)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(a.reps):
            net.residual_model.encode_symbols(xd, _jpeg=net.jpeg.decompress(bufs, "cuda").float().contiguous())
        torch.cuda.synchronize()
        t_sym = (time.perf_counter() - t0) / a.reps
    nbytes = sum(len(s) for grp in (c["strings"][0][0], c["strings"][0][1], c["strings"][1]) for s in grp)
    res = dict(tiles=a.tiles, h=a.h, w=a.w, mpixel=px / 1e6, enc_ms=t_enc * 1e3, dec_ms=t_dec * 1e3,
               enc_gpu_symbols_ms=t_sym * 1e3, enc_mpx_s=px / t_enc / 1e6, dec_mpx_s=px / t_dec / 1e6,
               encdec_mpx_s=px / (t_enc + t_dec) / 1e6, residual_bpp=8.0 * nbytes / px, host_cores=os.cpu_count(),
               x_hat_range=[float(d["x_hat"].min()), float(d["x_hat"].max())])
    print(json.dumps(res))


if __name__ == "__main__":
    main()
