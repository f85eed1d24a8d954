"""Single-call anatomy of compress / decompress with the device coder on the bench's model and image:
kernel time of the coder launches (CUDA events), escape fraction, mode fraction.
usage: python tools/time_codec_dev.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hyres_b200  # noqa: E402
from hyres_b200 import ops, synthetic  # noqa: E402


def main():
    torch.manual_seed(1926)
    net = hyres_b200.ResidualJPEGCompression(jpeg_quality=1)
    net.update(force=True)
    net = net.cuda().eval()
    codec = net.residual_model
    x = synthetic.synthetic_image(8, 704, 512, seed=7).cuda()
    with torch.no_grad():
        for coder in ("host", "device"):
            codec.coder = coder
            for _ in range(2):
                c = net.compress(x)
                net.decompress(c)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            c = net.compress(x)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            net.decompress(c)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            print(f"{coder}: compress {1e3 * (t1 - t0):.1f} ms, decompress {1e3 * (t2 - t1):.1f} ms")
        jd = net.jpeg.compress_device(x)[1]
        s = codec.encode_symbols(x, _jpeg=jd)
        gc, ebm = codec.gaussian_conditional, codec.entropy_bottleneck
        for k in ("slot_a", "slot_na"):
            sl = s[k]
            print(k, "escape fraction", float((sl < 0).float().mean()), "chunks with an escape",
                  float((sl.reshape(-1, 32) < 0).any(dim=1).float().mean()))
        print("zero symbols a / na:", float((s["sym_a"] == 0).float().mean()), float((s["sym_na"] == 0).float().mean()))
        ty, tz = gc.device_tables("cuda"), ebm.device_tables("cuda")
        groups = [(s["sym_z"], ebm.device_indexes(s["sym_z"].size(), "cuda"), tz, False),
                  (s["sym_a"], s["slot_a"], ty, True), (s["sym_na"], s["slot_na"], ty, True)]
        for name, g in (("z", groups[:1]), ("y both passes", groups[1:]), ("all", groups)):
            for _ in range(2):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                out = ops.rans_encode_device(g)
                torch.cuda.synchronize()
                ms = 1e3 * (time.perf_counter() - t0)
            print(f"device encode {name}: {ms:.1f} ms, {sum(len(b) for grp in out for b in grp)} bytes")


if __name__ == "__main__":
    main()
