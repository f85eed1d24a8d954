#!/usr/bin/env python
"""Fused residual-unit kernel (csrc/ru_fused.cu) against (a) a torch fp32 restatement that rounds
to bf16 where the kernel does and (b) the three-launch conv path; optional timing.

    python tools/check_ru.py [--perf]
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyres_b200 import ops  # noqa: E402
from hyres_b200.ops import ACT_NONE, ACT_RELU, EPI_ADD  # noqa: E402

CASES = [
    dict(name="1x16x8", B=1, H=16, W=8, relu=True),
    dict(name="1x32x48", B=1, H=32, W=48, relu=True),
    dict(name="2x40x24_ragged", B=2, H=40, W=24, relu=False),     # H not a multiple of the 16-row tile
    dict(name="3x24x20_ragged_w", B=3, H=24, W=20, relu=True),    # W not a multiple of the 8-col tile
    dict(name="2x128x192", B=2, H=128, W=192, relu=False),
]


def bf(x):
    return x.bfloat16().float()


def make_layers(seed):
    g = torch.Generator().manual_seed(seed)
    w1 = torch.randn(64, 128, 1, 1, generator=g) * (1.0 / 128 ** 0.5)
    w2 = torch.randn(64, 64, 3, 3, generator=g) * (1.0 / 576 ** 0.5)
    w3 = torch.randn(128, 64, 1, 1, generator=g) * (1.0 / 64 ** 0.5)
    b1, b2, b3 = (torch.randn(n, generator=g) * 0.2 for n in (64, 64, 128))
    c1 = ops.ConvLayer(w1, b1)
    c2 = ops.ConvLayer(w2, b2, pad=1)
    c3 = ops.ConvLayer(w3, b3)
    return (w1, b1, w2, b2, w3, b3), (c1, c2, c3)


def torch_ref(x_nhwc, W, relu):
    w1, b1, w2, b2, w3, b3 = [t.cuda() for t in W]
    x = x_nhwc.float().permute(0, 3, 1, 2)
    t1 = bf(F.relu(F.conv2d(x, bf(w1), b1)))
    t2 = bf(F.relu(F.conv2d(t1, bf(w2), b2, padding=1)))
    o = F.conv2d(t2, bf(w3), b3) + x
    if relu:
        o = F.relu(o)
    return o.permute(0, 2, 3, 1).contiguous()


def run_case(idx, perf=False):
    c = CASES[idx]
    W, (c1, c2, c3) = make_layers(100 + idx)
    g = torch.Generator().manual_seed(idx)
    x = torch.randn(c["B"], c["H"], c["W"], 128, generator=g).bfloat16().cuda()
    out = ops.ru_fused(x, c1, c2, c3, c["relu"])
    torch.cuda.synchronize()
    ref = torch_ref(x, W, c["relu"])
    a, _, _ = c1(x, act=ACT_RELU)
    b, _, _ = c2(a, act=ACT_RELU)
    un, _, _ = c3(b, epi=EPI_ADD, aux0=x, act=ACT_RELU if c["relu"] else ACT_NONE)
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    err_ref = float((out.float() - ref).abs().max()) / scale
    err_un = float((out.float() - un.float()).abs().max()) / scale
    r = dict(name=c["name"], err_vs_torch=err_ref, err_vs_unfused=err_un, ok=bool(err_ref < 1e-2 and err_un < 1e-2))
    if perf:
        for fn, key in ((lambda: ops.ru_fused(x, c1, c2, c3, c["relu"]), "fused_ms"),):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            r[key] = e0.elapsed_time(e1) / 10
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--perf", action="store_true")
    a = ap.parse_args()
    ok = True
    for i in range(len(CASES)):
        r = run_case(i)
        ok &= r["ok"]
        print(json.dumps(r))
    if a.perf:
        # the bench shape: 16 x 256 x 384 x 128 (g_a.3 / g_s.5 residual units)
        W, (c1, c2, c3) = make_layers(7)
        x = torch.randn(16, 256, 384, 128, device="cuda").bfloat16()
        out = torch.empty_like(x)
        for _ in range(3):
            ops.ru_fused(x, c1, c2, c3, True, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 20
        for _ in range(n):
            ops.ru_fused(x, c1, c2, c3, True, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        px = 16 * 256 * 384
        flops = 2.0 * px * (128 * 64 + 576 * 64 + 64 * 128)
        gb = px * 128 * 2 * 2 / 1e9
        print(json.dumps(dict(shape="16x256x384x128", ms=ms, tflops=flops / ms / 1e9, gbs=gb / (ms * 1e-3))))
        ref = torch_ref(x[:1], W, True)
        print(json.dumps(dict(big_err=float((out[:1].float() - ref).abs().max()) / float(ref.abs().max()))))
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
