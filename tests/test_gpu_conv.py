"""tcgen05 implicit-GEMM convolution (csrc/conv_tc.cu through hyres_conv_run) against a torch fp32
convolution of the same bf16-rounded operands.  Tolerances: fp32 outputs 2e-4 of the output
range (accumulation order only), bf16 outputs 1e-2 (one bf16 rounding of the result)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import check_conv  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("idx", range(len(check_conv.CASES)), ids=[c["name"] for c in check_conv.CASES])
def test_conv_case(build_lib, idx):
    case = dict(check_conv.CASES[idx])
    check_conv.CASES[idx] = {k: v for k, v in case.items() if k != "perf"}  # correctness only
    try:
        r = check_conv.run_case(idx)
    finally:
        check_conv.CASES[idx] = case
    assert r["ok"], r


def test_conv_argument_errors(build_lib):
    import torch
    from hyres_b200 import _lib, ops
    w = torch.randn(64, 64, 3, 3)
    with pytest.raises(_lib.HyresError):
        ops.ConvLayer(w, None, stride=3, pad=1)  # unsupported stride
    with pytest.raises(_lib.HyresError):
        ops.ConvLayer(torch.randn(64, 60, 1, 1), None)  # channels not a multiple of 8
    layer = ops.ConvLayer(w, None, stride=1, pad=1)
    with pytest.raises(ValueError):
        layer(torch.zeros(1, 8, 8, 32, dtype=torch.bfloat16, device="cuda"))  # wrong channel count
    with pytest.raises(ValueError):
        layer(torch.zeros(1, 8, 8, 64, dtype=torch.float32, device="cuda"))  # wrong dtype
    x = torch.zeros(1, 8, 8, 64, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(_lib.HyresError):
        layer(x, epi=ops.EPI_ADD)  # epilogue operand missing
    o, _, _ = layer(x)
    torch.cuda.synchronize()
    assert o.shape == (1, 8, 8, 64) and float(o.abs().max()) == 0.0


import check_ru  # noqa: E402


@pytest.mark.parametrize("idx", range(len(check_ru.CASES)), ids=[c["name"] for c in check_ru.CASES])
def test_fused_residual_unit(build_lib, idx):
    """csrc/ru_fused.cu (ResidualUnit / ResidualBottleneckBlock in one kernel) against the torch fp32
    restatement with bf16 storage of the two intermediates, and against the three-launch conv path;
    tolerance 1e-2 of the output range (one bf16 rounding of t1, t2 and the result)."""
    r = check_ru.run_case(idx)
    assert r["ok"], r


def test_fused_residual_unit_argument_errors(build_lib):
    import torch
    from hyres_b200 import _lib, ops
    c1 = ops.ConvLayer(torch.randn(64, 128, 1, 1), None)
    c2 = ops.ConvLayer(torch.randn(64, 64, 3, 3), None, pad=1)
    c3 = ops.ConvLayer(torch.randn(128, 64, 1, 1), None)
    bad = ops.ConvLayer(torch.randn(64, 64, 3, 3), None, pad=2, dil=2)
    assert ops.ru_supported(c1, c2, c3) and not ops.ru_supported(c1, bad, c3)
    x = torch.zeros(1, 16, 8, 128, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(_lib.HyresError):
        ops.ru_fused(x, c1, bad, c3, True)
    with pytest.raises(_lib.HyresError):
        ops.ru_fused(x, c1, c2, c3, True, out=x)  # in place


def _im2col_weight(w, kpad):
    """[N,3,K,K] -> the [N,kpad,1,1] GEMM weight over k = (r*K+s)*3 + c (engine.py ga0 / conv_in)."""
    import torch
    n, _, k, _ = w.shape
    w2 = torch.zeros(n, kpad, 1, 1)
    w2[:, :k * k * 3, 0, 0] = w.permute(0, 2, 3, 1).reshape(n, k * k * 3)
    return w2


@pytest.mark.parametrize("shape", [(2, 64, 96), (1, 40, 48), (3, 32, 160)], ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("variant", ["ga0_5x5s2", "conv_in_3x3"])
def test_conv3ch_fused_first_layer(build_lib, variant, shape):
    """csrc/conv_c3.cu: g_a.0 on x - jpeg and refine.conv_in (+PReLU) on jpeg + r_hat, im2col built on chip.
    src is fp32-exact; the activation matches a torch fp32 convolution of the bf16-rounded operands to
    1e-2 of its range (one bf16 rounding of the result); ragged tiles (W/2 = 24, H = 40) included."""
    import torch
    import torch.nn.functional as F
    from hyres_b200 import ops
    B, H, W = shape
    g = torch.Generator().manual_seed(H * 7 + W)
    a = torch.rand(B, 3, H, W, generator=g).cuda()
    b = torch.rand(B, 3, H, W, generator=g).cuda()
    if variant == "ga0_5x5s2":
        w = torch.randn(128, 3, 5, 5, generator=g) / 75 ** 0.5
        bias = torch.randn(128, generator=g) * 0.1
        layer = ops.ConvLayer(_im2col_weight(w, 128), bias)
        src, out = ops.conv3ch(layer, 5, 2, a, b, sign=-1)
        want_src = a - b
        ref = F.conv2d(want_src.bfloat16().float(), w.cuda().bfloat16().float(), bias.cuda(), stride=2, padding=2)
        alone, out2 = ops.conv3ch(layer, 5, 2, want_src)
    else:
        w = torch.randn(64, 3, 3, 3, generator=g) / 27 ** 0.5
        bias = torch.randn(64, generator=g) * 0.1
        layer = ops.ConvLayer(_im2col_weight(w, 64), bias)
        src, out = ops.conv3ch(layer, 3, 1, a, b, sign=1, act=ops.ACT_PRELU, slope=0.25)
        want_src = a + b
        ref = F.prelu(F.conv2d(want_src.bfloat16().float(), w.cuda().bfloat16().float(), bias.cuda(), padding=1),
                      torch.tensor([0.25], device="cuda"))
        alone, out2 = ops.conv3ch(layer, 3, 1, want_src, act=ops.ACT_PRELU, slope=0.25)
    torch.cuda.synchronize()
    assert torch.equal(src, want_src)
    assert alone is want_src and torch.equal(out2, out)  # single-operand form: same activation, no copy of src
    ref = ref.permute(0, 2, 3, 1)
    err = float((out.float() - ref).abs().max()) / float(ref.abs().max())
    assert err < 1e-2, err
