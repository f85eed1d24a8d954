#!/usr/bin/env python
"""Where a pipelined compress + decompress step spends its host time: wall-clock spans of the coder calls
(ctypes, GIL released), of the device-to-host symbol copies (which wait for the GPU) and of whole jobs, per step.

    python tools/trace_codec.py [--workers 4] [--steps 12]
"""
import argparse
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workers", type=int, default=4)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--coder", default="auto", choices=["auto", "host", "device"])
    a = ap.parse_args()
    import torch
    import hyres_b200
    from hyres_b200 import coder, entropy, synthetic
    spans = []
    lock = threading.Lock()

    def wrap(obj, name, tag):
        fn = getattr(obj, name)

        def inner(*args, **kw):
            t0 = time.perf_counter()
            try:
                return fn(*args, **kw)
            finally:
                with lock:
                    spans.append((tag, threading.get_ident(), t0, time.perf_counter()))
        setattr(obj, name, inner)

    wrap(coder, "encode_batch", "rans_encode")
    wrap(coder, "decode_batch", "rans_decode")
    orig = entropy.EntropyModel._host_i32.__func__

    def host_i32(cls, t, slot):
        t0 = time.perf_counter()
        try:
            return orig(cls, t, slot)
        finally:
            with lock:
                spans.append(("d2h_wait", threading.get_ident(), t0, time.perf_counter()))
    entropy.EntropyModel._host_i32 = classmethod(host_i32)

    torch.manual_seed(1926)
    net = hyres_b200.ResidualJPEGCompression(jpeg_quality=1)
    net.update(force=True)
    net = net.cuda().eval()
    net.residual_model.coder = a.coder
    from hyres_b200 import ops
    wrap(ops, "rans_encode_device", "dev_encode(incl. wait)")
    wrap(ops, "rans_upload", "dev_upload")
    wrap(ops, "read_small", "dev_decode_wait")
    wrap(net.jpeg, "decompress", "jpeg_decode")
    wrap(net.jpeg, "compress_device", "jpeg_encode")
    wrap(net, "compress", "compress_job")
    wrap(net, "decompress", "decompress_job")
    xs = [synthetic.synthetic_image(8, 704, 512, seed=7 + k).cuda() for k in range(4)]
    pipe = hyres_b200.CodecPipeline(net, workers=a.workers, reuse_host_buffers=True)
    with torch.no_grad():
        pipe.warm(xs[0])
        for _ in pipe.roundtrip((xs[i % 4] for i in range(a.workers + 2)), to_host=False):
            pass
        torch.cuda.synchronize()
        spans.clear()
        t0 = time.perf_counter()
        for _ in pipe.roundtrip((xs[i % 4] for i in range(a.steps)), to_host=False):
            pass
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
    pipe.close()
    print(f"workers {a.workers}: {1e3 * wall / a.steps:.2f} ms per step ({8 * 704 * 512 * a.steps / wall / 1e6:.1f} Mpixel/s), "
          f"{os.cpu_count()} host cores")
    agg = {}
    for tag, _, s, e in spans:
        d = agg.setdefault(tag, [0, 0.0])
        d[0] += 1
        d[1] += e - s
    for tag, (n, t) in sorted(agg.items()):
        print(f"  {tag:16s} {n / a.steps:5.1f} calls/step  {1e3 * t / n:7.2f} ms each  {1e3 * t / a.steps:7.2f} ms/step summed over threads")


if __name__ == "__main__":
    main()
