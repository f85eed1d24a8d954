"""The oracle (oracle/hyres_oracle.py) against fixtures produced by the reference's OWN in-tree
code (tests/golden/make_golden.py ran /root/reference/models/*.py unmodified; only the absent
compressai / turbojpeg packages were stood in for).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import load_golden


def _t(a):
    return torch.from_numpy(np.asarray(a))


def test_attention_block_matches_reference(oracle):
    g = load_golden("layers")
    att = oracle.AttentionBlock(32).eval()
    att.load_state_dict({k[len("att_sd."):]: _t(v) for k, v in g.items() if k.startswith("att_sd.")})
    with torch.no_grad():
        y = att(_t(g["att_x"]))
    torch.testing.assert_close(y, _t(g["att_y"]), rtol=1e-5, atol=1e-6)


def test_checkerboard_masked_conv_matches_reference(oracle):
    g = load_golden("layers")
    cm = oracle.CheckboardMaskedConv2d(8, 16, kernel_size=5, padding=2, stride=1).eval()
    with torch.no_grad():
        cm.weight.copy_(_t(g["cm_w_before"]))
        cm.bias.copy_(_t(g["cm_b"]))
        y = cm(_t(g["cm_x"]))
    assert torch.equal(cm.mask, _t(g["cm_mask"]))
    # 12 live taps, odd (i+j) parity, centre dead (models/layers/checkerboard.py:43-44)
    m = cm.mask[0, 0]
    assert int(m.sum()) == 12 and m[2, 2] == 0
    assert all(int(m[i, j]) == ((i + j) & 1) for i in range(5) for j in range(5))
    torch.testing.assert_close(y, _t(g["cm_y"]), rtol=1e-5, atol=1e-6)
    # Q4: the stored weights are masked in place by the call
    assert torch.equal(cm.weight.detach(), _t(g["cm_w_after"]))


def test_multiscale_refine_matches_reference(oracle):
    g = load_golden("layers")
    rf = oracle.MultiScaleRefine(3, 64).eval()
    rf.load_state_dict({k[len("rf_sd."):]: _t(v) for k, v in g.items() if k.startswith("rf_sd.")})
    with torch.no_grad():
        y = rf(_t(g["rf_x"]))
    torch.testing.assert_close(y, _t(g["rf_y"]), rtol=1e-4, atol=1e-6)


def test_quantizer_matches_reference(oracle):
    g = load_golden("layers")
    q = oracle.Quantizer()
    assert torch.equal(q.quantize(_t(g["q_x"]), "ste"), _t(g["q_ste"]))
    assert torch.equal(q.quantize(_t(g["q_x"]), "other"), _t(g["q_round"]))
    n = q.quantize(torch.zeros(1000), "noise")
    assert n.abs().max() <= 0.5 and n.std() > 0.2


@pytest.mark.parametrize("tag", ["codec64", "codec96x160"])
def test_codec_matches_reference_files(oracle, oracle_net, golden_weights_ok, tag):
    """Full LightWeightCheckerboard / ResidualJPEGCompression: oracle classes vs the reference's
    models/checkerboard.py + models/hyres.py run on the same weights and inputs."""
    if not golden_weights_ok:
        pytest.skip("regenerated weights differ from the fixture's (different torch build)")
    g = load_golden(tag)
    codec = oracle_net.residual_model
    residual = _t(g["residual"])
    with torch.no_grad():
        c = codec.compress(residual, return_intermediates=True)
        f = codec(residual)
        d = codec.decompress(c["strings"], c["shape"])
        w = oracle_net(_t(g["x"]), jpeg=(_t(g["jpeg_decoded"]), float(g["jpeg_bpp"])))
    tol = dict(rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(c["_y"], _t(g["y"]), **tol)
    torch.testing.assert_close(c["_z"], _t(g["z"]), **tol)
    torch.testing.assert_close(c["_anchor_params"], _t(g["params_a"]), **tol)
    torch.testing.assert_close(c["_non_anchor_params"], _t(g["params_na"]), **tol)
    exact = True
    for k in ("sym_z", "sym_a", "sym_na", "idx_a", "idx_na"):
        got, want = c["_" + k].int(), _t(g[k]).int()
        match = (got == want).float().mean().item()
        # threading changes oneDNN's summation order, so allow a tie flip in 10^4 elements
        assert match >= 0.9999, f"{k}: match {match}"
        exact &= match == 1.0
    if exact:
        assert c["strings"][0][0][0] == g["str_a"].tobytes()
        assert c["strings"][0][1][0] == g["str_na"].tobytes()
        assert c["strings"][1][0] == g["str_z"].tobytes()
    assert list(c["shape"]) == list(g["shape"])
    torch.testing.assert_close(f["x_hat"], _t(g["fwd_x_hat"]), rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(f["likelihoods"]["z"], _t(g["fwd_lik_z"]), rtol=1e-3, atol=1e-6)
    # likelihoods of y are discontinuous in the symbols: compare the rate instead of elements
    npx = residual.shape[0] * residual.shape[2] * residual.shape[3]
    bpp_y = (-f["likelihoods"]["y"].log2().sum() / npx).item()
    assert abs(bpp_y - float(g["bpp_y"])) < 2e-3 * max(1.0, float(g["bpp_y"]))
    torch.testing.assert_close(d["x_hat"], _t(g["dec_x_hat"]), rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(w["x_hat"], _t(g["w_x_hat"]), rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(w["residual_hat"], _t(g["w_residual_hat"]), rtol=1e-3, atol=1e-4)
    # Q3: decompress clamps the residual, forward does not
    assert d["x_hat"].min() >= 0 and d["x_hat"].max() <= 1
    assert f["x_hat"].min() < 0
    crit = oracle.RateDistortionLoss(lmbda=0.008)
    lo = crit(w, _t(g["x"]))
    assert abs(float(lo["mse_loss"]) - float(g["mse255"])) < 1e-3 * float(g["mse255"])
    assert abs(float(lo["z_bpp_loss"]) - float(g["bpp_z"])) < 1e-4


def test_cfg1_summary(oracle, oracle_net, golden_weights_ok):
    """BASELINE.json configs[0]: one synthetic 256x256 residual through the codec on CPU."""
    if not golden_weights_ok:
        pytest.skip("regenerated weights differ from the fixture's")
    g = load_golden("cfg1_summary")
    x = oracle.synthetic_residual(1, 256, 256)
    with torch.no_grad():
        f = oracle_net.residual_model(x)
    assert f["x_hat"].shape == (1, 3, 256, 256)  # models/checkerboard.py:290
    torch.testing.assert_close(f["x_hat"][0, :, 120:136, 120:136], _t(g["x_hat_center"]), rtol=1e-3, atol=1e-4)
    assert abs(f["x_hat"].double().abs().sum().item() - float(g["x_hat_abs_sum"])) < 1e-3 * float(g["x_hat_abs_sum"])
    assert abs(f["likelihoods"]["z"].double().log2().sum().item() - float(g["log2_lik_z"])) < 1e-3 * abs(float(g["log2_lik_z"]))
    assert abs(f["likelihoods"]["y"].double().log2().sum().item() - float(g["log2_lik_y"])) < 2e-3 * abs(float(g["log2_lik_y"]))


def test_param_counts(oracle_net):
    """assets/model_weights.png: 10.14 M parameters."""
    codec = sum(p.numel() for p in oracle_net.residual_model.parameters())
    refine = sum(p.numel() for p in oracle_net.refine.parameters())
    assert codec == 10_137_219
    assert refine == 238_061


def test_build_indexes_is_bucketize(oracle, oracle_net):
    gc = oracle_net.residual_model.gaussian_conditional
    s = torch.cat([torch.rand(1000) * 300, gc.scale_table, gc.scale_table * (1 + 1e-7), torch.tensor([0.0, -1.0, 0.11])])
    idx = gc.build_indexes(s)
    want = torch.bucketize(torch.clamp(s, min=0.11), gc.scale_table[:-1].contiguous(), right=False)
    assert torch.equal(idx.long(), want)


def test_cdf_tables_wellformed(oracle_net):
    for em in (oracle_net.residual_model.gaussian_conditional, oracle_net.residual_model.entropy_bottleneck):
        cdf, ln = em._quantized_cdf, em._cdf_length
        for i in range(cdf.size(0)):
            row = cdf[i, : ln[i]]
            assert row[0] == 0 and row[-1] == 65536
            assert (row[1:] > row[:-1]).all()
    gc = oracle_net.residual_model.gaussian_conditional
    assert tuple(gc._quantized_cdf.shape) == (64, 3133)
    assert int(gc._offset.min()) == -1565
