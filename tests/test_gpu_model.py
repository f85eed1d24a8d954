"""The drop-in model API on the GPU (LightWeightCheckerboard / ResidualJPEGCompression of
hyres_b200) against the CPU oracle on the same seeded weights and inputs.

Two arithmetic modes of the entropy-critical trunk (models.LightWeightCheckerboard.precision / codec_precision):
  "fp32h2" (compress / decompress default): compared with the oracle's fp32 mode -- the reference's semantics --
           y, z, entropy params to 5e-5 of their range, integer streams >= 0.999 equal end to end (every mismatch a
           numerical tie or its consequence: tests/test_gpu_precise.py has the stage-by-stage accounting);
  "bf16"   (forward default): activations stored in bf16 with fp32 accumulation, compared with the oracle's
           bf16-storage mode at x_hat / residual_hat 3e-2 of the output range (max abs), y, z, entropy params
           2e-2 / 5e-2 / 8e-2 of their range, rate 1 % relative, integer streams by match fraction.
g_s and MultiScaleRefine run in bf16 in both modes (x_hat: 3e-2 abs).  Encoder and decoder of the product are
bit-consistent with each other: decompress(compress(x)) reproduces forward(x) exactly when both use one mode."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nets(build_lib, oracle_net):
    import hyres_b200
    pnet = hyres_b200.ResidualJPEGCompression()
    pnet.load_state_dict(oracle_net.state_dict())
    return oracle_net, pnet.cuda().eval()


def _rel(got, want):
    return (got - want).abs().max().item() / max(want.abs().max().item(), 1e-12)


@pytest.mark.parametrize("B,H,W", [(1, 64, 64), (2, 96, 160), (1, 256, 256)])
def test_codec_stages_vs_oracle(nets, oracle, B, H, W):
    onet, pnet = nets
    x = oracle.synthetic_image(B, H, W, seed=9)
    jpeg_dec, _ = onet.jpeg(x)
    residual = x - jpeg_dec
    nchw = lambda t: t.permute(0, 3, 1, 2).float().cpu()  # noqa: E731
    # fp32-equivalent trunk (the codec default) against the reference's fp32 semantics
    assert pnet.residual_model.codec_precision == "fp32h2"
    with torch.no_grad():
        s = pnet.residual_model.encode_symbols(residual.cuda())
        with oracle.precision("fp32"):
            oc = onet.residual_model.compress(residual, return_intermediates=True)
    assert _rel(nchw(s["y"]), oc["_y"]) < 5e-5
    assert _rel(nchw(s["z"]), oc["_z"]) < 5e-5
    assert _rel(nchw(s["params_a"]), oc["_anchor_params"]) < 5e-5
    for k, lo in (("sym_z", 1.0), ("sym_a", 0.9999), ("idx_a", 0.9999), ("sym_na", 0.999), ("idx_na", 0.995)):
        got, want = s[k].cpu(), oc["_" + k].int()
        match = (got == want).float().mean().item()
        assert match >= lo, f"{k}: {match}"
        assert (got - want).abs().max().item() <= 1
    # plain bf16 trunk against the oracle's bf16-storage mode
    keep = pnet.residual_model.codec_precision
    pnet.residual_model.codec_precision = "bf16"
    try:
        with torch.no_grad():
            s = pnet.residual_model.encode_symbols(residual.cuda())
            with oracle.precision("bf16"):
                oc = onet.residual_model.compress(residual, return_intermediates=True)
    finally:
        pnet.residual_model.codec_precision = keep
    assert _rel(nchw(s["y"]), oc["_y"]) < 2e-2
    assert _rel(nchw(s["z"]), oc["_z"]) < 2e-2
    assert _rel(nchw(s["params_a"]), oc["_anchor_params"]) < 5e-2
    assert _rel(nchw(s["params_na"]), oc["_non_anchor_params"]) < 8e-2
    for k, lo in (("sym_z", 0.97), ("sym_a", 0.98), ("sym_na", 0.96), ("idx_a", 0.90), ("idx_na", 0.80)):
        got, want = s[k].cpu(), oc["_" + k].int()
        assert got.shape == want.shape
        match = (got == want).float().mean().item()
        assert match >= lo, f"{k}: {match}"
    for k in ("sym_z", "sym_a", "sym_na"):
        assert (s[k].cpu() - oc["_" + k].int()).abs().max().item() <= 2  # only tie flips


@pytest.mark.parametrize("B,H,W", [(1, 64, 96), (2, 128, 128)])
def test_forward_vs_oracle(nets, oracle, B, H, W):
    onet, pnet = nets
    x = oracle.synthetic_image(B, H, W, seed=10)
    jpeg = onet.jpeg(x)
    with torch.no_grad():
        pw = pnet(x.cuda(), jpeg=jpeg)
        with oracle.precision("bf16"):
            ow = onet(x, jpeg=jpeg)
    assert set(pw.keys()) == {"x_hat", "likelihoods", "jpeg_bpp_loss", "jpeg_decoded", "residual", "residual_hat"}
    assert torch.equal(pw["residual"].cpu(), ow["residual"])
    assert torch.equal(pw["jpeg_decoded"].cpu(), ow["jpeg_decoded"])
    assert pw["x_hat"].shape == x.shape and pw["x_hat"].dtype == torch.float32
    assert pw["likelihoods"]["y"].shape == (B, 192, H // 8, W // 8)
    assert pw["likelihoods"]["z"].shape == (B, 128, H // 32, W // 32)
    assert _rel(pw["residual_hat"].cpu(), ow["residual_hat"]) < 3e-2
    assert (pw["x_hat"].cpu() - ow["x_hat"]).abs().max().item() < 3e-2
    assert 0 <= pw["x_hat"].min() and pw["x_hat"].max() <= 1
    npx = B * H * W
    for k in ("y", "z"):
        bp = (-pw["likelihoods"][k].double().log2().sum() / npx).item()
        bo = (-ow["likelihoods"][k].double().log2().sum() / npx).item()
        assert abs(bp - bo) <= 1e-2 * bo, (k, bp, bo)
    import hyres_b200
    lp = hyres_b200.RateDistortionLoss(lmbda=0.008)(pw, x.cuda())
    with oracle.precision("bf16"):
        lo = oracle.RateDistortionLoss(lmbda=0.008)(ow, x)
    for k in ("loss", "bpp_loss", "mse_loss", "y_bpp_loss", "z_bpp_loss"):
        assert float(lp[k]) == pytest.approx(float(lo[k]), rel=1e-2), k
    # fused statistics path gives the same loss as the tensor path
    stats = torch.zeros(2, dtype=torch.float64, device="cuda")
    with torch.no_grad():
        pw2 = pnet(x.cuda(), jpeg=jpeg, stats=stats)
    lp2 = hyres_b200.RateDistortionLoss(lmbda=0.008)(pw2, x.cuda(), stats=stats)
    assert float(lp2["loss"]) == pytest.approx(float(lp["loss"]), rel=1e-6)
    assert torch.equal(pw2["x_hat"], pw["x_hat"])  # deterministic


def test_compress_decompress_roundtrip_and_self_consistency(nets, oracle):
    """models/checkerboard.py:167-240: the dict layout of the reference, and -- the property that
    makes the bitstream decodable -- the decoder recomputes exactly the encoder's parameters."""
    onet, pnet = nets
    codec = pnet.residual_model
    B, H, W = 2, 96, 160
    x = oracle.synthetic_image(B, H, W, seed=12)
    jpeg_dec, _ = onet.jpeg(x)
    residual = (x - jpeg_dec).cuda()
    with torch.no_grad():
        c = codec.compress(residual)
        assert set(c.keys()) == {"strings", "shape", "time"}
        assert isinstance(c["shape"], torch.Size) and tuple(c["shape"]) == (H // 32, W // 32)
        (sa, sna), sz = c["strings"]
        assert len(sa) == len(sna) == len(sz) == B and all(isinstance(t, bytes) for t in sa + sna + sz)
        d = codec.decompress(c["strings"], c["shape"])
        codec.precision = codec.codec_precision  # forward on the trunk the codec uses
        try:
            f = codec(residual)
        finally:
            codec.precision = "bf16"
    assert d["x_hat"].shape == (B, 3, H, W)
    assert torch.equal(d["x_hat"], f["x_hat"].clamp(0, 1))  # Q3, and encoder/decoder bit-consistency
    # the strings decode to exactly the symbols the encoder produced
    s = codec.encode_symbols(residual)
    gc = codec.gaussian_conditional
    assert torch.equal(gc.decode_symbols(sa, s["idx_a"]), s["sym_a"].cpu())
    assert torch.equal(gc.decode_symbols(sna, s["idx_na"]), s["sym_na"].cpu())
    # and are byte-identical to what the reference coder (oracle C restatement) makes of them
    for i in range(B):
        want = oracle.rans_encode_with_indexes(s["sym_a"][i].cpu().numpy().ravel(), s["idx_a"][i].cpu().numpy().ravel(),
                                               gc._quantized_cdf.cpu().numpy(), gc._cdf_length.cpu().numpy(),
                                               gc._offset.cpu().numpy())
        assert sa[i] == want
    inf = codec.inference(residual)
    assert torch.equal(inf["x_hat"], d["x_hat"]) and set(inf["time"]) == {"compression", "decompression", "total"}


def test_wrapper_compress_decompress(nets, oracle):
    onet, pnet = nets
    x = oracle.synthetic_image(1, 64, 96, seed=13)
    with torch.no_grad():
        c = pnet.compress(x.cuda())
        assert "jpeg_buffers" in c and len(c["jpeg_buffers"]) == 1
        d = pnet.decompress(c)
        with oracle.precision("fp32"):
            oc = onet.compress(x, jpeg_buffers=c["jpeg_buffers"])
            od = onet.decompress(oc)
    assert d["x_hat"].shape == x.shape and 0 <= d["x_hat"].min() and d["x_hat"].max() <= 1
    # PSNR of the product's reconstruction is the oracle's to 0.05 dB
    psnr = lambda t: (-10 * torch.log10((t - x).pow(2).mean())).item()  # noqa: E731
    assert abs(psnr(d["x_hat"].cpu()) - psnr(od["x_hat"])) < 0.05


def test_container_survives_decompress(nets, oracle):
    """decompress(unpack(pack(compress(x)))) == decompress(compress(x)) bit for bit."""
    from hyres_b200 import container
    onet, pnet = nets
    x = oracle.synthetic_image(2, 64, 96, seed=12)
    with torch.no_grad():
        c = pnet.compress(x.cuda())
        a = pnet.decompress(c)["x_hat"]
        b = pnet.decompress(container.unpack(container.pack(c)))["x_hat"]
    assert torch.equal(a, b)


def test_oracle_decodes_product_strings_when_symbols_agree(nets, oracle):
    """Cross-implementation decode: the oracle's decoder reads the product's hyper-latent string
    (z symbols depend only on the analysis trunk) whenever the z symbols agree."""
    onet, pnet = nets
    x = oracle.synthetic_residual(1, 64, 64, seed=21)
    with torch.no_grad():
        s = pnet.residual_model.encode_symbols(x.cuda())
        c = pnet.residual_model.compress(x.cuda())
        oeb = onet.residual_model.entropy_bottleneck
        z_hat = oeb.decompress(c["strings"][1], c["shape"])
    med = oeb._get_medians().detach().reshape(1, -1, 1, 1)
    assert torch.equal(z_hat, s["sym_z"].cpu().float() + med)


def test_noisequant_training_mode_runs(nets, oracle):
    onet, pnet = nets
    x = oracle.synthetic_residual(2, 64, 64, seed=22).cuda()
    codec = pnet.residual_model
    codec.train()
    try:
        with torch.no_grad():
            a = codec(x, noisequant=True)
            b = codec(x, noisequant=True)
    finally:
        codec.eval()
    assert a["x_hat"].shape == x.shape and torch.isfinite(a["x_hat"]).all()
    assert not torch.equal(a["likelihoods"]["y"], b["likelihoods"]["y"])  # fresh noise per call
    assert a["likelihoods"]["y"].min() >= 1e-9 and a["likelihoods"]["z"].min() >= 1e-9


def test_input_validation(nets):
    _, pnet = nets
    with pytest.raises(ValueError):
        pnet.residual_model(torch.zeros(1, 3, 60, 64, device="cuda"))  # not a multiple of 32
    with pytest.raises(ValueError):
        pnet.residual_model(torch.zeros(1, 1, 64, 64, device="cuda"))
    with pytest.raises(RuntimeError):
        pnet.residual_model(torch.zeros(1, 3, 64, 64))  # host tensor: no CPU fallback


def test_full_size_properties_cfg3_tile(nets, oracle):
    """BASELINE.json configs[2] at full size on one GPU: the 2048x1408 image as 8 tiles of
    512x704; size-independent properties: decode(encode) round trip, forward == decompress."""
    from hyres_b200 import dist as D
    onet, pnet = nets
    codec = pnet.residual_model
    x = oracle.synthetic_residual(1, 1408, 2048, seed=30)
    tiles = torch.cat([x[:, :, h0:h1, w0:w1] for h0, h1, w0, w1 in D.tile_grid(1408, 2048, 2, 4)], 0).cuda()
    assert tiles.shape == (8, 3, 704, 512)
    with torch.no_grad():
        c = codec.compress(tiles)
        d = codec.decompress(c["strings"], c["shape"])
        codec.precision = codec.codec_precision
        try:
            f = codec(tiles)
        finally:
            codec.precision = "bf16"
    assert torch.equal(d["x_hat"], f["x_hat"].clamp(0, 1))
    nbytes = sum(len(s) for grp in (c["strings"][0][0], c["strings"][0][1], c["strings"][1]) for s in grp)
    assert all(len(s) % 4 == 0 and len(s) >= 8 for s in c["strings"][1])
    assert 0 < nbytes * 8 / (1408 * 2048) < 64


def test_full_size_properties_cfg2(nets, oracle):
    """BASELINE.json configs[1] shape (batch 16 of 768x512): determinism and the checksum of
    checksums (fused sums == sums of the returned tensors); per-image independence."""
    import hyres_b200
    onet, pnet = nets
    x = oracle.synthetic_image(16, 512, 768, seed=31)
    jd = (x * 0.9 + 0.05)  # stand-in for the JPEG stage output (a boundary input)
    stats = torch.zeros(2, dtype=torch.float64, device="cuda")
    with torch.no_grad():
        a = pnet(x.cuda(), jpeg=(jd, 0.3), stats=stats)
        b = pnet(x.cuda(), jpeg=(jd, 0.3))
        one = pnet(x[5:6].cuda(), jpeg=(jd[5:6], 0.3))
    assert torch.equal(a["x_hat"], b["x_hat"]) and torch.equal(a["likelihoods"]["y"], b["likelihoods"]["y"])
    sy = a["likelihoods"]["y"].double().log2().sum().item()
    sz = a["likelihoods"]["z"].double().log2().sum().item()
    assert stats[0].item() == pytest.approx(sy, rel=1e-6) and stats[1].item() == pytest.approx(sz, rel=1e-6)
    assert torch.equal(one["x_hat"][0], a["x_hat"][5])  # images are independent: sharding by image is exact
    lo = hyres_b200.RateDistortionLoss(lmbda=0.008)(a, x.cuda())
    assert torch.isfinite(lo["loss"])


def test_host_pipeline_matches_direct_calls(nets, oracle):
    """hyres_b200.HostPipeline (pinned host batches, H2D on a copy stream under the previous batch's kernels, results
    read one batch late) returns, in order, exactly the scalars of direct forward + RateDistortionLoss calls."""
    import hyres_b200
    onet, pnet = nets
    crit = hyres_b200.RateDistortionLoss(lmbda=0.008)
    batches = []
    for k in range(5):
        x = oracle.synthetic_image(2, 64, 96, seed=40 + k).pin_memory()
        jd, bpp = onet.jpeg(x)
        batches.append((x, jd.contiguous().pin_memory(), bpp))
    want = []
    with torch.no_grad():
        for x, jd, bpp in batches:
            stats = torch.zeros(2, dtype=torch.float64, device="cuda")
            out = pnet(x.cuda(), jpeg=(jd.cuda(), bpp), stats=stats)
            lo = crit(out, x.cuda(), stats=stats)
            want.append({k: float(lo[k]) for k in ("loss", "bpp_loss", "mse_loss")})
    pipe = hyres_b200.HostPipeline(pnet, crit)
    got = list(pipe.run(iter(batches)))
    assert len(got) == len(want)
    for g, w in zip(got, want):
        for k in w:
            assert g[k] == pytest.approx(w[k], rel=1e-6), (k, g, w)
    assert pipe.h2d_bytes == 5 * 2 * 2 * 3 * 64 * 96 * 4 and pipe.d2h_bytes == 5 * 24
    assert list(pipe.run(iter([]))) == []


def test_host_pipeline_graph_replay_equals_eager_launches(nets, oracle):
    """Batches without an injected JPEG result run the device JPEG stage; after one eager pass per slot the pipeline
    replays a CUDA graph of forward + loss.  Results must equal the eager pipeline's to the last bit, the graphs must
    be dropped when a parameter changes, and a new batch shape must fall back to eager launches."""
    import hyres_b200
    _, pnet = nets
    crit = hyres_b200.RateDistortionLoss(lmbda=0.008)
    batches = [oracle.synthetic_image(2, 64, 96, seed=60 + k).pin_memory() for k in range(7)]
    eager = list(hyres_b200.HostPipeline(pnet, crit, use_graph=False).run(iter(batches)))
    pipe = hyres_b200.HostPipeline(pnet, crit)
    got = list(pipe.run(iter(batches)))
    assert got == eager
    assert all(s.graph is not None for s in pipe.slots), "no CUDA graph was captured"
    # direct calls agree too (device JPEG stage inside forward)
    with torch.no_grad():
        stats = torch.zeros(2, dtype=torch.float64, device="cuda")
        out = pnet(batches[3].cuda(), stats=stats)
        lo = crit(out, batches[3].cuda(), stats=stats)
    assert float(lo["loss"]) == got[3]["loss"]
    # a parameter update invalidates the captured graphs (the packed weights are refreshed on the eager path)
    p = pnet.refine.fusion[2].bias
    saved = p.detach().clone()
    with torch.no_grad():
        p.add_(0.25)
    try:
        changed = list(pipe.run(iter(batches[:3])))
        ref = list(hyres_b200.HostPipeline(pnet, crit, use_graph=False).run(iter(batches[:3])))
        assert changed == ref and changed[0]["loss"] != got[0]["loss"]
    finally:
        with torch.no_grad():
            p.copy_(saved)  # bit-exact restore for the tests that follow
    # another shape: slots are re-staged and run eagerly first
    other = [oracle.synthetic_image(1, 96, 64, seed=70 + k).pin_memory() for k in range(3)]
    assert list(pipe.run(iter(other))) == list(hyres_b200.HostPipeline(pnet, crit, use_graph=False).run(iter(other)))


def test_full_size_properties_cfg4_refine(nets, oracle):
    """BASELINE.json configs[3] shape (MultiScaleRefine on frozen-codec output, batch 32 of 512x512): the refine
    engine alone, fed x0 = jpeg + r_hat.  Size-independent properties: x0 is the exact fp32 sum; images are
    independent (any slice of the batch reproduces its rows bit for bit); the wrapper's reconstruction is the clamp
    of their sum.  Oracle parity is checked on one whole small image of the same generator (the three-scale
    network with its global SE pool has no crop-local receptive field), at the bf16 trunk's tolerance."""
    onet, pnet = nets
    eng = pnet.refine_engine()
    g = torch.Generator().manual_seed(44)
    jpeg = torch.rand(32, 3, 512, 512, generator=g).cuda()
    r_hat = (torch.randn(32, 3, 512, 512, generator=g) * 0.05).cuda()
    with torch.no_grad():
        x0, refined = eng(r_hat, jpeg)
        x0_b, refined_b = eng(r_hat[7:9].contiguous(), jpeg[7:9].contiguous())
        out = pnet._reconstruct(jpeg, r_hat)
    assert torch.equal(x0, jpeg + r_hat)
    assert refined.shape == (32, 3, 512, 512) and torch.isfinite(refined).all()
    assert torch.equal(refined[7:9], refined_b) and torch.equal(x0[7:9], x0_b)
    assert torch.equal(out, torch.clamp(x0 + refined, 0, 1))
    # oracle parity on one small image (bf16-storage restatement)
    js, rs = jpeg[:1, :, :96, :128].contiguous(), r_hat[:1, :, :96, :128].contiguous()
    with torch.no_grad():
        _, got = eng(rs, js)
        with oracle.precision("bf16"):
            want = onet.refine((js + rs).cpu())
    assert _rel(got.cpu(), want) < 3e-2


def test_codec_pipeline_matches_direct_calls(nets, oracle):
    """hyres_b200.CodecPipeline (several batches in flight on worker threads / CUDA streams) returns, in order,
    exactly the strings and reconstructions of back-to-back public compress() / decompress() calls, from pinned
    host batches and from device batches."""
    import hyres_b200
    _, pnet = nets
    batches = [oracle.synthetic_image(2, 64, 96, seed=80 + k).pin_memory() for k in range(7)]
    want = []
    with torch.no_grad():
        for x in batches:
            c = pnet.compress(x.cuda())
            want.append((c, pnet.decompress(c)["x_hat"].cpu()))
    pipe = hyres_b200.CodecPipeline(pnet, workers=3)
    try:
        got = list(pipe.roundtrip(iter(batches)))
        assert len(got) == len(want)
        for (c, x_hat), (wc, wx) in zip(got, want):
            assert c["strings"] == wc["strings"] and tuple(c["shape"]) == tuple(wc["shape"])
            assert [b.getvalue() for b in c["jpeg_buffers"]] == [b.getvalue() for b in wc["jpeg_buffers"]]
            assert not x_hat.is_cuda and x_hat.is_pinned() and torch.equal(x_hat, wx)
        assert pipe.h2d_bytes == 7 * 2 * 3 * 64 * 96 * 4 and pipe.d2h_bytes == pipe.h2d_bytes
        # device batches in, device reconstructions out; separate compress / decompress stages
        cs = list(pipe.compress(x.cuda() for x in batches[:3]))
        xs = list(pipe.decompress(iter(cs), to_host=False))
        for x_hat, (_, wx) in zip(xs, want[:3]):
            assert x_hat.is_cuda and torch.equal(x_hat.cpu(), wx)
        assert list(pipe.roundtrip(iter([]))) == []
    finally:
        pipe.close()


def test_export_and_load_prepacked_model(nets, oracle, tmp_path):
    """hyres_b200.export_model / load_exported (the src/updata.py:50-78 step): CDF tables in the state dict plus the
    packed device operands of every layer; the loaded model re-packs nothing and reproduces strings and
    reconstructions bit for bit.  A changed weight falls back to packing from the state dict."""
    import hyres_b200
    from hyres_b200 import ops
    _, pnet = nets
    x = oracle.synthetic_image(2, 64, 96, seed=90).cuda()
    path = tmp_path / "hyres_export.pt"
    blob = hyres_b200.export_model(pnet, str(path), update=False)
    assert blob["format"] == 1 and len(blob["packed"]) > 100
    assert "residual_model.gaussian_conditional._quantized_cdf" in blob["state_dict"]
    with torch.no_grad():
        c = pnet.compress(x)
        want = pnet.decompress(c)["x_hat"]
        before = dict(ops.PACKED_STATS)
        net2 = hyres_b200.load_exported(str(path))
        assert ops.PACKED_STATS["packed"] == before["packed"]  # nothing was packed from fp32 weights
        # (layers with identical weights -- the freshly initialised GDN gammas -- share one exported entry)
        assert ops.PACKED_STATS["imported"] - before["imported"] >= len(blob["packed"])
        c2 = net2.compress(x)
        assert c2["strings"] == c["strings"]
        assert torch.equal(net2.decompress(c2)["x_hat"], want)
        assert torch.equal(net2(x)["x_hat"], pnet(x)["x_hat"])
        # a checkpoint whose weights moved since the export: the stale operands are ignored for that layer
        blob["state_dict"]["refine.fusion.2.bias"] = blob["state_dict"]["refine.fusion.2.bias"] + 0.25
        before = dict(ops.PACKED_STATS)
        net3 = hyres_b200.load_exported(blob)
        assert ops.PACKED_STATS["packed"] - before["packed"] == 1
        assert not torch.equal(net3.decompress(c)["x_hat"], want)


def test_spatially_sharded_compress_is_byte_identical_to_whole_image(nets, oracle):
    """SURVEY 8e tier T-B: one image coded tile by tile with halos (as 4 emulated ranks and as one rank owning all
    tiles) gives the integers and the strings of the whole-image compress(), bit for bit; without the halo it does
    not (that is tier T-A: independent tiles)."""
    from hyres_b200 import spatial
    _, pnet = nets
    codec = pnet.residual_model
    x = oracle.synthetic_image(1, 768, 1024, seed=101).cuda()
    with torch.no_grad():
        whole = pnet.compress(x)
        jd, _ = pnet.jpeg.forward_device(x)
        res = x - jd
        ref = codec.encode_symbols(res)
        parts = [spatial.encode_symbols_sharded(codec, res, 2, 2, ranks=(r, 4)) for r in range(4)]
        for k in ("sym_z", "sym_a", "idx_a", "sym_na", "idx_na"):
            got = sum(p[k] for p in parts)  # what the all-reduce computes: every element has one writer
            assert torch.equal(got, ref[k]), k
        sharded = spatial.compress_sharded(pnet, x, rows=2, cols=2)
        assert sharded["strings"] == whole["strings"] and tuple(sharded["shape"]) == tuple(whole["shape"])
        assert [b.getvalue() for b in sharded["jpeg_buffers"]] == [b.getvalue() for b in whole["jpeg_buffers"]]
        assert torch.equal(pnet.decompress(sharded)["x_hat"], pnet.decompress(whole)["x_hat"])
        no_halo = spatial.encode_symbols_sharded(codec, res, 2, 2, halo=0, ranks=(0, 1))
        assert not torch.equal(no_halo["sym_a"], ref["sym_a"])
    wins = spatial.spatial_windows(1408, 2048, 2, 4)
    assert len(wins) == 8 and wins[0] == ((0, 704, 0, 512), (0, 960, 0, 768))
    assert all(v % 32 == 0 for t, w in wins for v in t + w)
