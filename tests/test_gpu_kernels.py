"""Memory-bound kernels of the hot path through the C-ABI, against (a) fixtures produced by the
reference's own files (tests/golden/*.npz) and (b) the CPU oracle on seeded inputs.

Bar: integer / byte outputs (symbols, CDF indexes, rANS strings) bit-exact; floating point within
the tolerance written at each assert."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _nhwc(a):
    return torch.from_numpy(np.asarray(a)).permute(0, 2, 3, 1).contiguous().cuda()


@pytest.fixture(scope="module")
def pcodec(build_lib, oracle_net):
    import hyres_b200
    net = hyres_b200.ResidualJPEGCompression()
    net.load_state_dict(oracle_net.state_dict())
    return net.cuda().eval().residual_model


@pytest.mark.parametrize("tag", ["codec64", "codec96x160"])
def test_symbols_indexes_strings_bit_exact_vs_reference_fixture(pcodec, golden_weights_ok, tag):
    """K8 + K13: fed the reference's y / params, the kernels give the reference's symbols and CDF
    indexes bit for bit, and the coder the reference's byte strings."""
    from hyres_b200 import ops
    g = load_golden(tag)
    y = _nhwc(g["y"])
    table = pcodec._scale_table("cuda")
    for ps, (p, sym, idx, s) in enumerate((("params_a", "sym_a", "idx_a", "str_a"),
                                           ("params_na", "sym_na", "idx_na", "str_na"))):
        prm = _nhwc(g[p])
        got_sym, got_idx, yq32, yq16 = ops.gc_symbols(y, prm, ps, table, 0.11)
        assert torch.equal(got_sym.cpu(), torch.from_numpy(g[sym].astype(np.int32)))
        assert torch.equal(got_idx.cpu(), torch.from_numpy(g[idx].astype(np.int32)))
        # decoder side: indexes from the scales alone, dequantised values = symbol + mean
        assert torch.equal(ops.gc_indexes(prm, table, pcodec.M, 0.11), got_idx)
        dq32, _ = ops.gc_dequant(got_sym, prm)
        assert torch.equal(dq32, yq32)
        M = pcodec.M
        want = got_sym.permute(0, 2, 3, 1).float() + prm[..., M:]
        assert torch.equal(yq32, want)
        # device front-end of the coder: slots (encoder) and codes (decoder) against their definition
        gcm = pcodec.gaussian_conditional
        rows = gcm.coder_rows("cuda")
        base, off, last = (r.long() for r in rows.cpu())
        r5 = ops.gc_symbols(y, prm, ps, table, 0.11, rows=rows)
        assert torch.equal(r5[0], got_sym) and torch.equal(r5[1], got_idx) and torch.equal(r5[2], yq32)
        ci, sv = got_idx.cpu().long(), got_sym.cpu().long()
        value = sv - off[ci]
        inside = (value >= 0) & (value < last[ci])
        assert torch.equal(r5[4].cpu().long(), torch.where(inside, base[ci] + value, -(ci + 1)))
        codes = ops.gc_codes(prm, ps, table, M := pcodec.M, rows, 0.11).cpu().long()
        hh, ww = ci.shape[-2:]
        anchor = ((torch.arange(hh).view(-1, 1) + torch.arange(ww).view(1, -1)) % 2 == 0).expand_as(ci)
        structural = anchor != (ps == 0)
        mu = prm[..., M:].permute(0, 3, 1, 2).cpu()
        known_val = torch.round(-mu).long() - off[ci]
        known = structural & (known_val >= 0) & (known_val < last[ci])
        assert torch.equal(codes, torch.where(known, (1 << 30) | (base[ci] + known_val), ci))
        assert torch.equal(sv[structural], torch.round(-mu).long()[structural])  # Q1: what the encoder coded there
        junk = torch.where(structural.cuda(), torch.full_like(got_sym, 777), got_sym)
        dq_pass, _ = ops.gc_dequant(junk, prm, pass_id=ps)
        assert torch.equal(dq_pass, yq32)  # structural symbols are recomputed, not read
        if golden_weights_ok:  # the CDF tables come from the regenerated weights / scale table
            strings = pcodec.gaussian_conditional.encode_symbols(got_sym, got_idx)
            assert strings[0] == g[s].tobytes()
            assert gcm.encode_symbol_groups([(got_sym, r5[4])], slots=True)[0][0] == g[s].tobytes()
            back = pcodec.gaussian_conditional.decode_symbols(strings, got_idx)
            assert torch.equal(back, got_sym.cpu())
            back2 = gcm.decode_symbols(strings, ops.gc_codes(prm, ps, table, M, rows, 0.11), codes=True)
            assert torch.equal(back2[~known], got_sym.cpu()[~known])
    if golden_weights_ok:
        z = _nhwc(g["z"])
        ebp, med = pcodec.engine().eb_params()
        eb = ops.eb_forward(z, ebp, med, want_symbols=True)
        assert torch.equal(eb["symbols"].cpu(), torch.from_numpy(g["sym_z"].astype(np.int32)))
        ebm = pcodec.entropy_bottleneck
        zs = ebm.encode_symbols(eb["symbols"], ebm._build_indexes(eb["symbols"].size()))
        assert zs[0] == g["str_z"].tobytes()
        lik_z = eb["lik"].cpu()
        torch.testing.assert_close(lik_z, torch.from_numpy(g["fwd_lik_z"]), rtol=1e-3, atol=2e-7)


@pytest.mark.parametrize("tag", ["codec64", "codec96x160"])
def test_likelihood_and_quantiser_vs_reference_fixture(pcodec, tag):
    """K6 + K7: STE quantisation of both passes (Q1: every position of both zero-filled tensors),
    y_hat, and the erfc likelihood under summed parameters (Q2).  fp32: rtol 1e-3, atol 2e-7
    (the difference of two fp32 erfc values cancels to ~1e-7 on both sides)."""
    from hyres_b200 import ops
    g = load_golden(tag)
    y, pa, pna = _nhwc(g["y"]), _nhwc(g["params_a"]), _nhwc(g["params_na"])
    M = pcodec.M
    yqa32, yqa16 = ops.gc_quant_pass(y, pa, 0)
    yqna32, _ = ops.gc_quant_pass(y, pna, 1)
    B, h, w, _ = y.shape
    ii, jj = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    anchor = (((ii + jj) & 1) == 0).cuda()[None, :, :, None]
    ya = torch.where(anchor, y, torch.zeros_like(y))
    yna = torch.where(anchor, torch.zeros_like(y), y)
    assert torch.equal(yqa32, torch.round(ya - pa[..., M:]) + pa[..., M:])
    assert torch.equal(yqna32, torch.round(yna - pna[..., M:]) + pna[..., M:])
    assert torch.equal(yqa16, yqa32.bfloat16())
    acc = torch.zeros(1, dtype=torch.float64, device="cuda")
    y_hat16, lik = ops.gc_merge_likelihood(y, pa, pna, yqa32, yqna32, sum_log2=acc)
    assert torch.equal(y_hat16, (yqa32 + yqna32).bfloat16())
    want = torch.from_numpy(g["fwd_lik_y"])
    torch.testing.assert_close(lik.cpu(), want, rtol=1e-3, atol=2e-7)
    assert lik.min() >= 1e-9
    ref_sum = want.double().log2().sum().item()
    assert abs(acc.item() - ref_sum) <= 1e-4 * abs(ref_sum)
    acc2 = torch.zeros(1, dtype=torch.float64, device="cuda")
    ops.reduce_log2(lik, acc2)
    assert abs(acc2.item() - lik.double().log2().sum().item()) <= 1e-6 * abs(ref_sum)  # fp32 log2f per element


def test_noise_quantiser_statistics(pcodec):
    from hyres_b200 import ops
    y = torch.randn(2, 16, 24, 192, device="cuda") * 3
    a, _ = ops.gc_quant_pass(y, None, 0, noise=True, seed=11)
    b, _ = ops.gc_quant_pass(y, None, 0, noise=True, seed=11)
    c, _ = ops.gc_quant_pass(y, None, 0, noise=True, seed=12)
    assert torch.equal(a, b) and not torch.equal(a, c)  # seeded, reproducible
    ii, jj = torch.meshgrid(torch.arange(16), torch.arange(24), indexing="ij")
    anchor = (((ii + jj) & 1) == 0).cuda()[None, :, :, None]
    ya = torch.where(anchor, y, torch.zeros_like(y))
    n = a - ya  # models/utils/quantization.py:6-10: x + U(-1/2, 1/2) on the zero-filled tensor
    assert n.abs().max() <= 0.5
    assert abs(n.mean().item()) < 5e-3 and abs(n.var().item() - 1 / 12) < 5e-3


def test_residual_addback_clamp_exact(build_lib):
    """K5 (models/hyres.py:48,62,66-67): fp32 elementwise, bit-exact (the residual and the add-back are produced by
    the fused first-layer kernels, conv_c3.cu -- tests/test_gpu_conv.py::test_conv3ch_fused_first_layer -- and by
    the split-precision im2col, tests/test_gpu_precise.py)."""
    from hyres_b200 import ops
    g = torch.Generator().manual_seed(2)
    x = torch.rand(2, 3, 64, 96, generator=g).cuda()
    j = torch.rand(2, 3, 64, 96, generator=g).cuda()
    r_hat = torch.randn(2, 3, 64, 96, generator=g).cuda() * 0.1
    x0 = j + r_hat
    refined = torch.randn(2, 3, 64, 96, generator=g).cuda()
    assert torch.equal(ops.final_clamp(x0, refined), torch.clamp(x0 + refined, 0, 1))
    acc = torch.zeros(1, dtype=torch.float64, device="cuda")
    ops.reduce_sqdiff(x0, x, acc)
    want = (x0.double() - x.double()).pow(2).sum().item()
    assert abs(acc.item() - want) <= 1e-6 * want


def test_entropy_bottleneck_vs_oracle(pcodec, oracle_net):
    """K9 (compressai EntropyBottleneck via models/checkerboard.py:96-101): eval-mode outputs are
    round(z - median) + median exactly; likelihoods rtol 1e-3 / atol 2e-7."""
    from hyres_b200 import ops
    oeb = oracle_net.residual_model.entropy_bottleneck
    g = torch.Generator().manual_seed(4)
    z = torch.randn(2, 128, 6, 10, generator=g) * 4
    with torch.no_grad():
        z_hat, lik = oeb(z, training=False)
    ebp, med = pcodec.engine().eb_params()
    r = ops.eb_forward(z.permute(0, 2, 3, 1).contiguous().cuda(), ebp, med, want_zhat_nchw=True, want_symbols=True)
    assert torch.equal(r["zhat_nchw"].cpu(), z_hat)
    torch.testing.assert_close(r["lik"].cpu(), lik, rtol=1e-3, atol=2e-7)
    medv = oeb._get_medians().detach().reshape(1, -1, 1, 1)
    assert torch.equal(r["symbols"].cpu(), torch.round(z - medv).int())
    assert torch.equal(ops.eb_dequant(r["symbols"], med), r["zhat_bf16"])


def test_refine_memory_ops_vs_oracle(build_lib, oracle_net):
    """K10 (models/layers/enhancement.py:15-21,36-40,96-108) on bf16-stored features."""
    from hyres_b200 import ops
    import torch.nn.functional as F
    rf = oracle_net.refine
    g = torch.Generator().manual_seed(6)
    feat = torch.randn(2, 64, 32, 48, generator=g).bfloat16()
    f16 = feat.permute(0, 2, 3, 1).contiguous().cuda()
    fc1, fc2 = rf.se_block.fc[0].weight.detach().cuda().contiguous(), rf.se_block.fc[2].weight.detach().cuda().contiguous()
    fs, fh, fq, pooled = ops.refine_se_scale_down(f16, fc1, fc2)
    with torch.no_grad():
        x = feat.float()
        want_pool = x.mean(dim=(2, 3))
        se = rf.se_block(x).bfloat16().float()  # stored bf16
        half = F.interpolate(se, scale_factor=0.5, mode="bilinear", align_corners=False)
        quarter = F.interpolate(se, scale_factor=0.25, mode="bilinear", align_corners=False)
    torch.testing.assert_close(pooled.cpu(), want_pool, rtol=1e-4, atol=1e-5)
    nchw = lambda t: t.float().permute(0, 3, 1, 2).cpu()  # noqa: E731
    torch.testing.assert_close(nchw(fs), se, rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(nchw(fh), half, rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(nchw(fq), quarter, rtol=2e-2, atol=2e-2)
    # upsample + concat + channel mean/max statistics + 7x7 attention
    f1 = torch.randn(2, 32, 64, 64, generator=g).bfloat16().cuda()
    f2 = torch.randn(2, 16, 32, 64, generator=g).bfloat16().cuda()
    f3 = torch.randn(2, 8, 16, 64, generator=g).bfloat16().cuda()
    pad1 = lambda t: F.pad(nchw(t), (1, 1, 1, 1), mode="replicate").permute(0, 2, 3, 1).bfloat16().contiguous().cuda()  # noqa: E731
    stats = ops.refine_stats3_tc(f1, pad1(f2), pad1(f3))
    with torch.no_grad():
        up2 = F.interpolate(nchw(f2), size=(32, 64), mode="bilinear", align_corners=False)
        up3 = F.interpolate(nchw(f3), size=(32, 64), mode="bilinear", align_corners=False)
        cat = torch.cat([nchw(f1), up2, up3], 1).bfloat16().float()
        want_att = rf.spatial_att(cat)[:, 0]
    torch.testing.assert_close(stats[..., 0].cpu(), cat.mean(1), rtol=1e-3, atol=2e-3)
    torch.testing.assert_close(stats[..., 1].cpu(), cat.max(1)[0], rtol=2e-2, atol=2e-2)
    w7 = rf.spatial_att.conv.weight.detach().reshape(-1).cuda().contiguous()
    att = ops.refine_spatial_att(stats, w7)
    torch.testing.assert_close(att.cpu(), want_att, rtol=1e-2, atol=1e-2)


def test_layout_helpers_exact(build_lib):
    from hyres_b200 import ops
    x = torch.randn(2, 24, 10, 14).cuda()
    n16 = ops.nchw_f32_to_nhwc_bf16(x)
    assert torch.equal(n16, x.permute(0, 2, 3, 1).bfloat16())
    assert torch.equal(ops.nhwc_bf16_to_nchw_f32(n16), n16.float().permute(0, 3, 1, 2))
    n32 = x.permute(0, 2, 3, 1).contiguous()
    assert torch.equal(ops.nhwc_to_nchw_f32(n32), x)
