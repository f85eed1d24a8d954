"""Training step (BASELINE.json configs[4]; src/utils/engine.py:29-90) on the GPU against the CPU oracle's autograd:
same weights, image, injected JPEG stage and injected noise tensors.

Stated tolerance: the product's convolutions (forward, data gradient, weight gradient) run on bf16 tensor cores with
bf16 activations / gradients between layers and fp32 master weights, the oracle in fp32.  Loss components agree to
1 %; per-parameter gradients to a cosine similarity >= 0.98 in the noise-quantiser mode (>= 0.999 weighted by
gradient norm); in the straight-through mode rounding flips between the two precisions lower the hyper-network's
gradients' agreement (>= 0.8, weighted >= 0.99); 20 Adam steps follow the oracle's loss curve within 5 %."""
import json
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import check_train  # noqa: E402

pytestmark = pytest.mark.gpu


def _fresh(oracle):
    import hyres_b200
    onet = oracle.make_model(seed=1926, wrapper=True, lively=True)
    pnet = hyres_b200.ResidualJPEGCompression()
    pnet.load_state_dict(onet.state_dict())
    return onet, pnet.cuda()


@pytest.mark.parametrize("noisequant", [True, False], ids=["noise", "ste"])
def test_gradients_vs_oracle_autograd(build_lib, oracle, noisequant):
    torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))
    onet, pnet = _fresh(oracle)
    x = oracle.synthetic_image(2, 64, 64, seed=3)
    r = check_train.grad_report(pnet, onet, oracle, x, noisequant=noisequant)
    print(json.dumps({k: r[k] for k in ("losses", "weighted_cos", "median_cos", "min_cos", "missing")}))
    assert r["missing"] == []  # every parameter the oracle differentiates gets a gradient, and no other
    for k, (got, want) in r["losses"].items():
        assert got == pytest.approx(want, rel=1e-2), k
    if noisequant:
        assert r["weighted_cos"] >= 0.999 and r["median_cos"] >= 0.999 and r["min_cos"] >= 0.98
    else:
        assert r["weighted_cos"] >= 0.99 and r["median_cos"] >= 0.995 and r["min_cos"] >= 0.8
    for name, v in r["params"].items():
        if v["oracle_norm"] > 1e-4:
            assert 0.8 < v["norm_ratio"] < 1.25, (name, v)


def test_twenty_adam_steps_follow_the_oracle(build_lib, oracle):
    torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))
    onet, pnet = _fresh(oracle)
    x = oracle.synthetic_image(2, 64, 64, seed=3)
    got, want = check_train.loss_curves(pnet, onet, oracle, x, steps=20)
    for k, (g, w) in enumerate(zip(got, want)):
        tol = 0.02 if k < 10 else 0.05
        assert g["loss"] == pytest.approx(w["loss"], rel=tol), (k, g["loss"], w["loss"])
        assert g["bpp_loss"] == pytest.approx(w["bpp_loss"], rel=tol), k
        assert g["aux_loss"] == pytest.approx(w["aux_loss"], rel=1e-4), k
    assert got[-1]["loss"] < 0.3 * got[0]["loss"]
    # the inference path sees the trained weights (packed layers are refreshed on the next call)
    pnet.eval()
    with torch.no_grad():
        out = pnet(x.cuda())
    assert torch.isfinite(out["x_hat"]).all()


CONV_GEOMS = [
    dict(name="1x1_128_64_relu", kind=0, cin=128, cout=64, k=1, relu=True),
    dict(name="3x3_64_64_relu", kind=0, cin=64, cout=64, k=3, relu=True),
    dict(name="3x3_dil2_64_64", kind=0, cin=64, cout=64, k=3, dil=2),
    dict(name="5x5_s2_128_192", kind=0, cin=128, cout=192, k=5, stride=2),
    dict(name="deconv_192_128", kind=1, cin=192, cout=128, k=5),
    dict(name="masked_5x5_192_384", kind=0, cin=192, cout=384, k=5, mask=True),
    dict(name="1x1_768_640_f32", kind=0, cin=768, cout=640, k=1, f32=True),
    dict(name="3x3_96_96", kind=0, cin=96, cout=96, k=3),
    dict(name="deconv_128_8_f32", kind=1, cin=128, cout=8, k=5, f32=True),
    dict(name="3x3_8_64", kind=0, cin=8, cout=64, k=3),
]


@pytest.mark.parametrize("c", CONV_GEOMS, ids=[c["name"] for c in CONV_GEOMS])
def test_conv_node_gradients_vs_torch(build_lib, c):
    """One convolution node: forward, data gradient (the tcgen05 kernel on the transposed problem) and weight / bias
    gradient against torch autograd of the same convolution on the bf16-rounded operands (fp32 math): outputs and
    gradients within 1.5e-2 of their range (one bf16 rounding of the result and of the incoming gradient)."""
    import torch.nn.functional as F
    from hyres_b200 import train as T
    g = torch.Generator().manual_seed(11)
    B, H, W = 2, 24, 40
    k, stride, dil = c["k"], c.get("stride", 1), c.get("dil", 1)
    cin, cout = c["cin"], c["cout"]
    wshape = (cin, cout, k, k) if c["kind"] == 1 else (cout, cin, k, k)
    w = (torch.randn(wshape, generator=g) / (cin * k * k) ** 0.5).cuda().requires_grad_()
    b = (torch.randn(cout, generator=g) * 0.1).cuda().requires_grad_()
    x = torch.randn(B, H, W, cin, generator=g).cuda().bfloat16().requires_grad_()
    mask = None
    if c.get("mask"):
        mask = torch.zeros(k, k, dtype=torch.uint8)
        mask[0::2, 1::2] = 1
        mask[1::2, 0::2] = 1
    pad = 2 if c["kind"] == 1 else dil * (k - 1) // 2
    tc = T.TrainConv(c["kind"], wshape, 2 if c["kind"] == 1 else stride, pad, dil, mask)
    from hyres_b200 import ops
    assert ops.wgrad_supported(tc) and T.wgrad_native_active()  # the weight gradient below is csrc/wgrad.cu's
    y = T.conv(x, w, b, tc, c.get("relu", False), c.get("f32", False))
    go = torch.randn(y.shape, generator=g).cuda().to(y.dtype)
    y.backward(go)
    got = (y.detach().float(), x.grad.float(), w.grad.clone(), b.grad.clone())
    # reference
    xr = x.detach().float().permute(0, 3, 1, 2).requires_grad_()
    wr = w.detach().bfloat16().float()
    if mask is not None:
        wr = wr * mask.cuda().float()  # models/layers/checkerboard.py:47 masks weight.data, not the gradient
    wr.requires_grad_()
    br = b.detach().clone().requires_grad_()
    if c["kind"] == 1:
        yr = F.conv_transpose2d(xr, wr, br, stride=2, padding=2, output_padding=1)
    else:
        yr = F.conv2d(xr, wr, br, stride=stride, padding=pad, dilation=dil)
    if c.get("relu"):
        yr = torch.relu(yr)
    yr.backward(go.float().bfloat16().float().permute(0, 3, 1, 2))
    want = (yr.detach().permute(0, 2, 3, 1), xr.grad.permute(0, 2, 3, 1), wr.grad, br.grad)
    for name, a, r in zip(("y", "dx", "dW", "db"), got, want):
        err = float((a - r).abs().max()) / max(float(r.abs().max()), 1e-12)
        assert err < 1.5e-2, (name, err)
