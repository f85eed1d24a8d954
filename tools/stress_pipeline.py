#!/usr/bin/env python
"""Stress the multi-stream codec pipeline: `mode` in compress / decompress / roundtrip, `--iters` batches with
`--in-flight` worker threads.  Exit code 0 = no CUDA fault.  (Diagnostic for gpurun.)"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hyres_b200  # noqa: E402
from hyres_b200 import synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["compress", "decompress", "roundtrip"])
    ap.add_argument("--iters", type=int, default=60)
    ap.add_argument("--in-flight", type=int, default=4)
    ap.add_argument("--tiles", type=int, default=8)
    a = ap.parse_args()
    torch.manual_seed(1926)
    net = hyres_b200.ResidualJPEGCompression(jpeg_quality=1)
    net.update(force=True)
    net = net.cuda().eval()
    xs = [synthetic.synthetic_image(a.tiles, 704, 512, seed=7 + k).cuda() for k in range(3)]
    pipe = hyres_b200.CodecPipeline(net, workers=a.in_flight, reuse_host_buffers=True)
    with torch.no_grad():
        cs = [net.compress(x) for x in xs]
        torch.cuda.synchronize()
        n = 0
        if a.mode == "compress":
            for c in pipe.compress(xs[i % 3] for i in range(a.iters)):
                n += 1
        elif a.mode == "decompress":
            for x_hat in pipe.decompress((cs[i % 3] for i in range(a.iters)), to_host=False):
                n += 1
        else:
            for c, x_hat in pipe.roundtrip((xs[i % 3] for i in range(a.iters)), to_host=False):
                n += 1
        torch.cuda.synchronize()
    pipe.close()
    print("ok", a.mode, n)


if __name__ == "__main__":
    main()
