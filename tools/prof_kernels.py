#!/usr/bin/env python
"""Launch one hot layer at its bench shape a few times (target of `ncu --set full`, or a quick timer).

    python tools/prof_kernels.py --case gate|gdn|c3x3|c1x1|ru|head0|head1|ctx|deconv|s2|gs8|fus0|p_ru1|p_ru2|p_ru3|p_s2 [--n 3] [--time] [--code 18]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyres_b200 import ops  # noqa: E402
from hyres_b200.ops import (ACT_NONE, ACT_PRELU, ACT_RELU, EPI_ADD, EPI_GATE, EPI_GDN, EPI_LINEAR,  # noqa: E402
                            EPI_PIXSCALE, HYRES_CONV, HYRES_DECONV_K5S2)

B = 16
MT = 0
CODE = 18  # nsplit code of the p_* cases: 18 = two half parts (fp32h2), 3 = three bf16 parts


def rnd(*shape):
    return torch.randn(*shape, device="cuda").bfloat16()


def make(case):
    g = torch.Generator().manual_seed(5)

    def w(co, ci, k):
        return torch.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5

    def bias(n):
        return torch.randn(n, generator=g) * 0.1

    if case == "gate":
        L = ops.ConvLayer(w(128, 128, 1), bias(128))
        x, a0, a1 = rnd(B, 256, 384, 128), rnd(B, 256, 384, 128), rnd(B, 256, 384, 128)
        out = torch.empty_like(x)
        return lambda: L(x, epi=EPI_GATE, aux0=a0, aux1=a1, out_bf16=out), 2.0 * B * 256 * 384 * 128 * 128, 4 * x.numel() * 2
    if case == "gate192":
        L = ops.ConvLayer(w(192, 192, 1), bias(192))
        x, a0, a1 = rnd(B, 64, 96, 192), rnd(B, 64, 96, 192), rnd(B, 64, 96, 192)
        out = torch.empty_like(x)
        return lambda: L(x, epi=EPI_GATE, aux0=a0, aux1=a1, out_bf16=out), 2.0 * B * 64 * 96 * 192 * 192, 4 * x.numel() * 2
    if case == "gdn":
        L = ops.ConvLayer(w(128, 128, 1).abs(), bias(128).abs() + 0.5)
        x = rnd(B, 256, 384, 128)
        out = torch.empty_like(x)
        return lambda: L(x, epi=EPI_GDN, aux0=x, x0_square=True, out_bf16=out), 2.0 * B * 256 * 384 * 128 * 128, 2 * x.numel() * 2
    if case == "c3x3":
        L = ops.ConvLayer(w(64, 64, 3), bias(64), pad=1)
        x = rnd(B, 512, 768, 64)
        out = torch.empty_like(x)
        return lambda: L(x, act=ACT_PRELU, slope=0.25, out_bf16=out), 2.0 * B * 512 * 768 * 64 * 576, 2 * x.numel() * 2
    if case == "c1x1":
        L = ops.ConvLayer(w(64, 128, 1), bias(64))
        x = rnd(B, 256, 384, 128)
        out = torch.empty(B, 256, 384, 64, device="cuda", dtype=torch.bfloat16)
        return lambda: L(x, act=ACT_RELU, out_bf16=out), 2.0 * B * 256 * 384 * 64 * 128, (x.numel() + out.numel()) * 2
    if case == "fus0":
        L = ops.ConvLayer(w(64, 192, 1), bias(64))
        x = rnd(B, 512, 768, 192)
        ps = torch.rand(B, 512, 768, device="cuda")
        out = torch.empty(B, 512, 768, 64, device="cuda", dtype=torch.bfloat16)
        return (lambda: L(x, epi=EPI_PIXSCALE, pixscale=ps, act=ACT_PRELU, slope=0.2, out_bf16=out),
                2.0 * B * 512 * 768 * 64 * 192, (x.numel() + out.numel()) * 2)
    if case in ("fus0up", "fus0ps", "fus0lin"):
        # MultiScaleRefine fusion[0] without the concat: 64 -> 64 1x1 + tensor-core up-add + pixel scale
        L = ops.ConvLayer(w(64, 64, 1), bias(64))
        x = rnd(B, 512, 768, 64)
        t2, t3 = rnd(B, 258, 386, 64), rnd(B, 130, 194, 64)
        ps = torch.rand(B, 512, 768, device="cuda")
        out = torch.empty(B, 512, 768, 64, device="cuda", dtype=torch.bfloat16)
        nbytes = (x.numel() + out.numel()) * 2
        if case == "fus0up":
            return (lambda: L(x, epi=EPI_PIXSCALE, pixscale=ps, act=ACT_PRELU, slope=0.2, out_bf16=out, up_t2=t2, up_t3=t3),
                    2.0 * B * 512 * 768 * 64 * 160, nbytes + (t2.numel() + t3.numel()) * 2 + ps.numel() * 4)
        if case == "fus0ps":
            return (lambda: L(x, epi=EPI_PIXSCALE, pixscale=ps, act=ACT_PRELU, slope=0.2, out_bf16=out),
                    2.0 * B * 512 * 768 * 64 * 64, nbytes + ps.numel() * 4)
        return lambda: L(x, act=ACT_PRELU, slope=0.2, out_bf16=out), 2.0 * B * 512 * 768 * 64 * 64, nbytes
    if case == "stats3":
        f1, s2p, s3p = rnd(B, 512, 768, 64), rnd(B, 258, 386, 64), rnd(B, 130, 194, 64)
        return (lambda: ops.refine_stats3_tc(f1, s2p, s3p), 2.0 * B * 512 * 768 * 64 * 96,
                (f1.numel() + s2p.numel() + s3p.numel()) * 2 + B * 512 * 768 * 8)
    if case == "jpeg":
        x = torch.rand(B, 3, 512, 768, device="cuda")
        return lambda: ops.jpeg_forward(x, 1), 0.0, x.numel() * 8 + B * 512 * 768 * 10
    if case == "ru":
        c1 = ops.ConvLayer(w(64, 128, 1), bias(64))
        c2 = ops.ConvLayer(w(64, 64, 3), bias(64), pad=1)
        c3 = ops.ConvLayer(w(128, 64, 1), bias(128))
        x = rnd(B, 256, 384, 128)
        out = torch.empty_like(x)
        return (lambda: ops.ru_fused(x, c1, c2, c3, True, out=out),
                2.0 * B * 256 * 384 * (128 * 64 + 576 * 64 + 64 * 128), 2 * x.numel() * 2)
    if case == "c3x3_96":
        L = ops.ConvLayer(w(96, 96, 3), bias(96), pad=1)
        x = rnd(B, 64, 96, 96)
        out = torch.empty_like(x)
        return lambda: L(x, act=ACT_RELU, out_bf16=out, mt=MT), 2.0 * B * 64 * 96 * 96 * 864, 2 * x.numel() * 2
    if case == "ru192":
        c1 = ops.ConvLayer(w(96, 192, 1), bias(96))
        c2 = ops.ConvLayer(w(96, 96, 3), bias(96), pad=1)
        c3 = ops.ConvLayer(w(192, 96, 1), bias(192))
        x = rnd(B, 64, 96, 192)
        out = torch.empty_like(x)

        def run():
            a, _, _ = c1(x, act=ACT_RELU)
            b, _, _ = c2(a, act=ACT_RELU)
            c3(b, epi=EPI_ADD, aux0=x, act=ACT_RELU, out_bf16=out)
        return run, 2.0 * B * 64 * 96 * (192 * 96 + 864 * 96 + 96 * 192), 2 * x.numel() * 2
    if case in ("head0", "head1", "head2"):
        ci, co = dict(head0=(768, 640), head1=(640, 512), head2=(512, 384))[case]
        L = ops.ConvLayer(w(co, ci, 1), bias(co), cin0=ci // 2 if case == "head0" else ci, cin1=ci // 2 if case == "head0" else 0)
        x = rnd(B, 64, 96, ci // 2 if case == "head0" else ci)
        x1 = rnd(B, 64, 96, ci // 2) if case == "head0" else None
        if case == "head2":
            o32 = torch.empty(B, 64, 96, co, device="cuda")
            return lambda: L(x, out_bf16=False, out_f32=o32, mt=MT), 2.0 * B * 64 * 96 * ci * co, x.numel() * 2 + o32.numel() * 4
        out = torch.empty(B, 64, 96, co, device="cuda", dtype=torch.bfloat16)
        return lambda: L(x, x1, act=ACT_RELU, out_bf16=out, mt=MT), 2.0 * B * 64 * 96 * ci * co, (B * 64 * 96 * ci + out.numel()) * 2
    if case == "ctx":
        mask = torch.zeros(5, 5, dtype=torch.uint8)
        mask[0::2, 1::2] = 1
        mask[1::2, 0::2] = 1
        L = ops.ConvLayer(w(384, 192, 5), bias(384), pad=2, tap_mask=mask)
        x = rnd(B, 64, 96, 192)
        out = torch.empty(B, 64, 96, 384, device="cuda", dtype=torch.bfloat16)
        return lambda: L(x, out_bf16=out, mt=MT), 2.0 * B * 64 * 96 * 192 * 12 * 384, (x.numel() + out.numel()) * 2
    if case == "deconv":
        wt = torch.randn(128, 128, 5, 5, generator=g) / (128 * 25 / 4) ** 0.5
        L = ops.ConvLayer(wt, bias(128), kind=HYRES_DECONV_K5S2)
        x = rnd(B, 128, 192, 128)
        out = torch.empty(B, 256, 384, 128, device="cuda", dtype=torch.bfloat16)
        return lambda: L(x, out_bf16=out, mt=MT), 2.0 * B * 128 * 192 * 128 * 25 * 128, (x.numel() + out.numel()) * 2
    if case == "s2":
        L = ops.ConvLayer(w(128, 128, 5), bias(128), stride=2, pad=2)
        x = rnd(B, 256, 384, 128)
        out = torch.empty(B, 128, 192, 128, device="cuda", dtype=torch.bfloat16)
        return lambda: L(x, out_bf16=out, mt=MT), 2.0 * B * 128 * 192 * 128 * 25 * 128, (x.numel() + out.numel()) * 2
    if case == "fus2":
        L = ops.ConvLayer(w(3, 64, 3), bias(3), pad=1)
        x = rnd(B, 512, 768, 64)
        o32 = torch.empty(B, 3, 512, 768, device="cuda")
        return (lambda: L(x, out_bf16=False, out_f32=o32.permute(0, 2, 3, 1)), 2.0 * B * 512 * 768 * 64 * 9 * 3,
                x.numel() * 2 + o32.numel() * 4)
    if case == "gs8":
        wt = torch.randn(128, 3, 5, 5, generator=g) / (128 * 25 / 4) ** 0.5
        L = ops.ConvLayer(wt, bias(3), kind=HYRES_DECONV_K5S2)
        x = rnd(B, 256, 384, 128)
        o32 = torch.empty(B, 3, 512, 768, device="cuda")
        return (lambda: L(x, out_bf16=False, out_f32=o32.permute(0, 2, 3, 1)), 2.0 * B * 256 * 384 * 128 * 25 * 3,
                x.numel() * 2 + o32.numel() * 4)
    if case in ("c3in", "c3ga"):
        a = torch.rand(B, 3, 512, 768, device="cuda")
        b2 = torch.rand(B, 3, 512, 768, device="cuda")
        if case == "c3in":
            w2 = torch.zeros(64, 64, 1, 1)
            w2[:, :27] = torch.randn(64, 27, 1, 1, generator=g) / 27 ** 0.5
            L = ops.ConvLayer(w2, bias(64))
            out = torch.empty(B, 512, 768, 64, device="cuda", dtype=torch.bfloat16)
            return (lambda: ops.conv3ch(L, 3, 1, a, b2, sign=1, act=ACT_PRELU, slope=0.2, out=out),
                    2.0 * B * 512 * 768 * 64 * 27, a.numel() * 12 + out.numel() * 2)
        w2 = torch.zeros(128, 128, 1, 1)
        w2[:, :75] = torch.randn(128, 75, 1, 1, generator=g) / 75 ** 0.5
        L = ops.ConvLayer(w2, bias(128))
        out = torch.empty(B, 256, 384, 128, device="cuda", dtype=torch.bfloat16)
        return (lambda: ops.conv3ch(L, 5, 2, a, b2, sign=-1, out=out),
                2.0 * B * 256 * 384 * 128 * 75, a.numel() * 12 + out.numel() * 2)
    if case.startswith("p_"):
        # split-precision (fp32-equivalent) layers of g_a at the codec bench shape: 8 tiles of 704x512 -> 8 x 352 x 256
        # positions.  p_ru1 / p_ru2 / p_ru3: the three layers of a ResidualUnit; p_gdn: conv -> x^2 parts, gamma GEMM
        code = CODE
        P = ops.split_parts(code)
        Bc, Hc, Wc = 8, 352, 256
        pos = Bc * Hc * Wc

        def parts(c):
            _, sp = ops.split_f32(torch.randn(Bc, Hc, Wc, c, device="cuda"), nsplit=code)
            return sp
        if case == "p_ru1":
            L = ops.ConvLayer(w(64, 128, 1), bias(64), nsplit=code)
            x = parts(128)
            return (lambda: L(x, act=ACT_RELU, out_bf16=False, out_split=True), 2.0 * pos * 128 * 64,
                    pos * (128 + 64) * 2 * P)
        if case == "p_ru2":
            L = ops.ConvLayer(w(64, 64, 3), bias(64), pad=1, nsplit=code)
            x = parts(64)
            return (lambda: L(x, act=ACT_RELU, out_bf16=False, out_split=True), 2.0 * pos * 576 * 64,
                    pos * (64 + 64) * 2 * P)
        if case == "p_ru3":
            L = ops.ConvLayer(w(128, 64, 1), bias(128), nsplit=code)
            x = parts(64)
            skip = torch.randn(Bc, Hc, Wc, 128, device="cuda")
            return (lambda: L(x, out_bf16=False, out_f32="nhwc", split_mode=ops.SPLIT_ADD, aux0_f32=skip, out_split=True),
                    2.0 * pos * 64 * 128, pos * (64 * 2 * P + 128 * 4 * 2 + 128 * 2 * P))
        if case in ("p_ru3_nof32", "p_ru3_noparts", "p_ru3_noskip"):  # sensitivity of p_ru3 to each of its streams
            L = ops.ConvLayer(w(128, 64, 1), bias(128), nsplit=code)
            x = parts(64)
            skip = torch.randn(Bc, Hc, Wc, 128, device="cuda")
            kw = dict(out_bf16=False, out_f32="nhwc", split_mode=ops.SPLIT_ADD, aux0_f32=skip, out_split=True)
            if case == "p_ru3_nof32":
                kw["out_f32"] = None
            elif case == "p_ru3_noparts":
                kw["out_split"] = None
            else:
                kw["split_mode"], kw["aux0_f32"] = ops.SPLIT_COPY, None
            return (lambda: L(x, **kw), 2.0 * pos * 64 * 128, pos * (64 * 2 * P + 128 * 4 * 2 + 128 * 2 * P))
        if case == "p_s2":
            L = ops.ConvLayer(w(128, 128, 5), bias(128), stride=2, pad=2, nsplit=code)
            x = parts(128)
            return (lambda: L(x, out_bf16=False, out_f32="nhwc", out_split=True, split_square=True, out_code=3),
                    2.0 * pos / 4 * 128 * 25 * 128, pos * 128 * 2 * P + pos // 4 * 128 * (4 + 6))
    raise SystemExit(f"unknown case {case}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", required=True)
    ap.add_argument("--n", type=int, default=3)
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--mt", type=int, default=0)
    ap.add_argument("--code", type=int, default=18)
    a = ap.parse_args()
    global MT, CODE
    MT = a.mt
    CODE = a.code
    for case in a.case.split(","):
        fn, flops, byts = make(case)
        for _ in range(a.n):
            fn()
        torch.cuda.synchronize()
        if a.time:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print(json.dumps(dict(case=case, ms=round(ms, 4), tflops=round(flops / ms / 1e9, 1), gbs=round(byts / ms / 1e6, 0))))


if __name__ == "__main__":
    main()
