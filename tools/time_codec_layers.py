#!/usr/bin/env python
"""Per-layer device time of ONE compress + decompress of the bench's 8-tile image (CUDA events around every
convolution launch, host coder), every layer geometry with its precision, sorted by time; plus the non-convolution
remainder (total GPU time of the two calls measured on the stream minus the convolution launches)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hyres_b200  # noqa: E402
from hyres_b200 import ops, synthetic  # noqa: E402


def main():
    torch.manual_seed(1926)
    net = hyres_b200.ResidualJPEGCompression(jpeg_quality=1)
    net.update(force=True)
    net = net.cuda().eval()
    net.residual_model.coder = "host"
    x = synthetic.synthetic_image(8, 704, 512, seed=7).cuda()
    with torch.no_grad():
        for _ in range(2):
            c = net.compress(x)
            net.decompress(c)
        torch.cuda.synchronize()
        for name, fn in (("compress", lambda: net.compress(x)), ("decompress", lambda: net.decompress(c))):
            ops.ConvLayer.profile_begin()
            fn()
            conv_ms, conv_n = ops.ConvLayer.profile_end()
            agg = {}
            for r in ops.ConvLayer.last_profile:
                key = (r["kind"], r["cin"], r["cout"], r["k"], r["stride"], r["OH"], r["OW"], r.get("nsplit", 1), r.get("split_mode", 0))
                t = agg.setdefault(key, [0, 0.0, 0.0])
                t[0] += 1
                t[1] += r["ms"]
                t[2] += 2.0 * r.get("alg_macs", 0) * r.get("products", 1)
            print(f"{name}: {conv_n} tensor-core launches, {conv_ms:.2f} ms")
            for key, (n, t, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                print(f"   kind {key[0]} {key[1]:4d}->{key[2]:4d} k{key[3]} s{key[4]} at {key[5]}x{key[6]} nsplit {key[7]} mode {key[8]}: "
                      f"{n:2d} launches {t:7.3f} ms  {fl / (t * 1e-3) / 1e12:7.1f} TF/s executed")


if __name__ == "__main__":
    main()
