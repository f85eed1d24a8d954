#!/usr/bin/env python
"""Time the memory-bound MultiScaleRefine kernels at the bench shape (16 x 512 x 768)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyres_b200 import ops  # noqa: E402

B, H, W = 16, 512, 768


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    f1 = torch.randn(B, H, W, 64, device="cuda").bfloat16()
    f2p = torch.randn(B, H // 2 + 2, W // 2 + 2, 64, device="cuda").bfloat16()
    f3p = torch.randn(B, H // 4 + 2, W // 4 + 2, 64, device="cuda").bfloat16()
    ms = timeit(lambda: ops.refine_stats3_tc(f1, f2p, f3p))
    gb = B * H * W * (128 + 32 + 8 + 8) / 1e9
    print(f"stats3_tc {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s")
    feat = torch.randn(B, H, W, 64, device="cuda").bfloat16()
    fc1, fc2 = torch.randn(4, 64, device="cuda") * 0.1, torch.randn(64, 4, device="cuda") * 0.1
    ms = timeit(lambda: ops.refine_se_scale_down(feat, fc1, fc2))
    gb = B * H * W * (128 * 2 + 128 + 32 + 8) / 1e9
    print(f"se_pool + se_scale_down {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s")
    stats = torch.randn(B, H, W, 2, device="cuda")
    w7 = torch.randn(98, device="cuda") * 0.1
    ms = timeit(lambda: ops.refine_spatial_att(stats, w7))
    print(f"spatial_att {ms:.3f} ms")


if __name__ == "__main__":
    main()
