"""MultiScaleRefine's fusion stage without the 192-channel concat (models/layers/enhancement.py:99-110):
``fusion[0](cat([s1, up2(s2), up4(s3)]) * att)`` is computed as ``att * (W1 s1 + up2(W2 s2) + up4(W3 s3)) + b`` with the
bilinear up-sampling done on the tensor cores inside the full-resolution layer.  Checked against plain PyTorch fp32
(``F.interpolate(..., mode="bilinear", align_corners=False)``) on the same bf16-rounded operands; tolerance 2e-2
absolute on O(1) outputs = bf16 output rounding plus the bf16 storage of the low-resolution products."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _nchw(t):
    return t.float().permute(0, 3, 1, 2)


@pytest.mark.parametrize("B,H,W", [(1, 32, 32), (2, 64, 96), (1, 128, 64)])
def test_upadd_conv_matches_interpolate_then_conv(build_lib, B, H, W):
    from hyres_b200 import ops
    from hyres_b200.ops import ACT_PRELU, EPI_PIXSCALE, HYRES_CONV
    g = torch.Generator().manual_seed(H * 7 + W)
    w = torch.randn(64, 192, 1, 1, generator=g) / 192 ** 0.5
    bias = torch.randn(64, generator=g) * 0.1
    f1 = torch.randn(B, H, W, 64, generator=g).bfloat16().cuda()
    f2 = torch.randn(B, H // 2, W // 2, 64, generator=g).bfloat16().cuda()
    f3 = torch.randn(B, H // 4, W // 4, 64, generator=g).bfloat16().cuda()
    att = torch.rand(B, H, W, generator=g).cuda()
    l1 = ops.ConvLayer(w[:, :64], bias, kind=HYRES_CONV)
    l2 = ops.ConvLayer(w[:, 64:128], None, kind=HYRES_CONV)
    l3 = ops.ConvLayer(w[:, 128:], None, kind=HYRES_CONV)
    t2, _, _ = l2(f2, out_pad=1)
    t3, _, _ = l3(f3, out_pad=1)
    ops.replicate_border(t2)
    ops.replicate_border(t3)
    # the padded tensors: interior = the 1x1 product, border = nearest interior pixel
    wb = w.bfloat16().float().cuda()
    ref2 = F.conv2d(_nchw(f2), wb[:, 64:128])
    ref3 = F.conv2d(_nchw(f3), wb[:, 128:])
    assert torch.allclose(_nchw(t2[:, 1:-1, 1:-1]), ref2, atol=2e-2, rtol=1e-2)
    assert torch.equal(t2[:, 0, 1:-1], t2[:, 1, 1:-1]) and torch.equal(t2[:, -1, 1:-1], t2[:, -2, 1:-1])
    assert torch.equal(t2[:, :, 0], t2[:, :, 1]) and torch.equal(t2[:, :, -1], t2[:, :, -2])
    assert torch.equal(t3[:, 0, 0], t3[:, 1, 1]) and torch.equal(t3[:, -1, -1], t3[:, -2, -2])
    h, _, _ = l1(f1, epi=EPI_PIXSCALE, pixscale=att, act=ACT_PRELU, slope=0.2, up_t2=t2, up_t3=t3)
    # reference on the stored (bf16) low-resolution products, so only the final rounding differs
    s2 = _nchw(t2[:, 1:-1, 1:-1])
    s3 = _nchw(t3[:, 1:-1, 1:-1])
    acc = (F.conv2d(_nchw(f1), wb[:, :64]) + F.interpolate(s2, scale_factor=2, mode="bilinear", align_corners=False)
           + F.interpolate(s3, scale_factor=4, mode="bilinear", align_corners=False))
    ref = F.prelu(acc * att[:, None] + bias.cuda()[None, :, None, None], torch.tensor([0.2], device="cuda"))
    err = (_nchw(h) - ref).abs().max().item()
    assert err < 2e-2, err
    # and against the reference formulation: conv over the concat of up-sampled branches
    cat = torch.cat([_nchw(f1), F.interpolate(_nchw(f2), scale_factor=2, mode="bilinear", align_corners=False),
                     F.interpolate(_nchw(f3), scale_factor=4, mode="bilinear", align_corners=False)], 1)
    ref_cat = F.prelu(F.conv2d(cat * att[:, None], wb, bias.cuda()), torch.tensor([0.2], device="cuda"))
    assert (_nchw(h) - ref_cat).abs().max().item() < 4e-2


def test_upadd_rejects_layers_it_cannot_run(build_lib):
    from hyres_b200 import ops
    from hyres_b200._lib import HyresError
    l = ops.ConvLayer(torch.randn(64, 64, 1, 1), None)
    x = torch.zeros(1, 48, 48, 64, dtype=torch.bfloat16, device="cuda")  # 48 is not a multiple of 32
    t2 = torch.zeros(1, 26, 26, 64, dtype=torch.bfloat16, device="cuda")
    t3 = torch.zeros(1, 14, 14, 64, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(HyresError):
        l(x, up_t2=t2, up_t3=t3)
    with pytest.raises(ValueError):
        l(x, up_t2=t2)


@pytest.mark.parametrize("B,H,W", [(1, 32, 32), (2, 64, 96)])
def test_tensor_core_stats_match_the_elementwise_kernel(build_lib, B, H, W):
    """hyres_refine_stats3_tc (bilinear up-sampling as GEMMs, channel fold in the epilogue) against PyTorch's
    bilinear interpolation of the concat: mean within 2e-3 (fp32 summation order), max within one bf16 ulp of O(1)
    values (2e-2)."""
    from hyres_b200 import ops
    g = torch.Generator().manual_seed(11)
    f1 = torch.randn(B, H, W, 64, generator=g).bfloat16().cuda()
    f2 = torch.randn(B, H // 2, W // 2, 64, generator=g).bfloat16().cuda()
    f3 = torch.randn(B, H // 4, W // 4, 64, generator=g).bfloat16().cuda()
    f2p = F.pad(_nchw(f2), (1, 1, 1, 1), mode="replicate").permute(0, 2, 3, 1).bfloat16().contiguous()
    f3p = F.pad(_nchw(f3), (1, 1, 1, 1), mode="replicate").permute(0, 2, 3, 1).bfloat16().contiguous()
    b = ops.refine_stats3_tc(f1, f2p, f3p)
    cat = torch.cat([_nchw(f1), F.interpolate(_nchw(f2), scale_factor=2, mode="bilinear", align_corners=False),
                     F.interpolate(_nchw(f3), scale_factor=4, mode="bilinear", align_corners=False)], 1)
    assert torch.allclose(b[..., 0], cat.mean(1), atol=2e-3)
    assert torch.allclose(b[..., 1], cat.amax(1), atol=2e-2)
