"""Host side of the drop-in boundary (no GPU): constructor / attribute / state-dict compatibility
with the reference API (models/checkerboard.py:24-283, models/hyres.py:9-181), CDF-table building,
and the loud failure when there is no CUDA device."""
import pytest
import torch


@pytest.fixture(scope="module")
def pnet(build_lib, oracle_net):
    import hyres_b200
    torch.manual_seed(0)
    net = hyres_b200.ResidualJPEGCompression()
    net.load_state_dict(oracle_net.state_dict())
    return net.eval()


def test_state_dict_keys_match_reference_layout(pnet, oracle_net):
    mine, theirs = pnet.state_dict(), oracle_net.state_dict()
    assert list(mine.keys()) == list(theirs.keys())
    for k in mine:
        assert mine[k].shape == theirs[k].shape, k
        assert torch.equal(mine[k].cpu(), theirs[k].cpu()), k
    # the key families SURVEY.md section 8b lists
    for k in ("residual_model.g_a.0.weight", "residual_model.g_a.1.beta", "residual_model.g_a.1.gamma_reparam.pedestal",
              "residual_model.g_a.2.conv1.weight", "residual_model.g_a.3.conv_a.0.conv.2.weight",
              "residual_model.g_a.3.conv_b.3.bias", "residual_model.context_prediction.mask",
              "residual_model.param_aggregation.4.weight", "residual_model.entropy_bottleneck.quantiles",
              "residual_model.entropy_bottleneck.matrices.0", "residual_model.entropy_bottleneck._quantized_cdf",
              "residual_model.gaussian_conditional.scale_table", "residual_model.gaussian_conditional._offset",
              "refine.conv_in.weight", "refine.se_block.fc.0.weight", "refine.scale2.2.bias",
              "refine.spatial_att.conv.weight", "refine.fusion.2.weight"):
        assert k in mine, k


def test_attributes_and_signatures(pnet):
    import inspect
    import hyres_b200
    c = pnet.residual_model
    for a in ("N", "M", "entropy_bottleneck", "gaussian_conditional", "quantizer", "g_a", "g_s", "h_a", "h_s",
              "context_prediction", "param_aggregation"):
        assert hasattr(c, a), a
    assert (c.N, c.M) == (128, 192)
    for a in ("jpeg", "residual_model", "refine"):
        assert hasattr(pnet, a)
    sig = inspect.signature(hyres_b200.LightWeightCheckerboard.forward)
    assert list(sig.parameters)[:3] == ["self", "x", "noisequant"] and sig.parameters["noisequant"].default is False
    sig = inspect.signature(hyres_b200.ResidualJPEGCompression.__init__)
    assert [sig.parameters[k].default for k in ("base_model", "jpeg_quality", "se_reduction")] == [None, 1, 1]
    assert list(inspect.signature(hyres_b200.LightWeightCheckerboard.decompress).parameters) == ["self", "strings", "shape"]
    assert list(inspect.signature(hyres_b200.ResidualJPEGCompression.decompress).parameters) == ["self", "compressed_data"]
    assert sum(p.numel() for p in c.parameters()) == 10_137_219


def test_update_builds_the_oracles_tables(build_lib, oracle):
    import hyres_b200
    torch.manual_seed(5)
    mine = hyres_b200.LightWeightCheckerboard()
    theirs = oracle.LightWeightCheckerboard()
    theirs.load_state_dict(mine.state_dict())
    assert mine.update(force=True) is True
    assert mine.update() is False  # tables exist: no rebuild without force (compressai semantics)
    theirs.update(force=True)
    for em in ("gaussian_conditional", "entropy_bottleneck"):
        for buf in ("_quantized_cdf", "_cdf_length", "_offset"):
            assert torch.equal(getattr(getattr(mine, em), buf), getattr(getattr(theirs, em), buf)), (em, buf)
    assert torch.equal(mine.gaussian_conditional.scale_table, hyres_b200.get_scale_table())
    assert float(mine.aux_loss()) == pytest.approx(float(theirs.aux_loss()), rel=1e-6)


def test_load_state_dict_resizes_tables_and_accepts_legacy_keys(build_lib, oracle_net):
    import hyres_b200
    sd = {k[len("residual_model."):]: v for k, v in oracle_net.state_dict().items() if k.startswith("residual_model.")}
    legacy = {}
    for k, v in sd.items():  # compressai <= 1.1 spelling
        k = k.replace("entropy_bottleneck.matrices.", "entropy_bottleneck._matrix")
        k = k.replace("entropy_bottleneck.biases.", "entropy_bottleneck._bias")
        k = k.replace("entropy_bottleneck.factors.", "entropy_bottleneck._factor")
        legacy[k] = v
    assert any("_matrix0" in k for k in legacy)
    net = hyres_b200.LightWeightCheckerboard.from_state_dict(legacy)
    assert tuple(net.gaussian_conditional._quantized_cdf.shape) == (64, 3133)
    assert torch.equal(net.entropy_bottleneck.matrices[2], sd["entropy_bottleneck.matrices.2"])


def test_wrapper_load_state_dict_split(build_lib, oracle_net):
    import hyres_b200
    sd = dict(oracle_net.state_dict())
    net = hyres_b200.ResidualJPEGCompression.from_state_dict(sd, jpeg_quality=7)
    assert net.jpeg.quality == 7
    assert torch.equal(net.refine.fusion[2].weight, sd["refine.fusion.2.weight"])
    sd["se_block.fc.0.weight"] = torch.zeros(1)
    with pytest.raises(AttributeError):  # models/hyres.py:159: self.se_block does not exist
        hyres_b200.ResidualJPEGCompression().load_state_dict(sd)


def test_no_cpu_fallback(pnet):
    """The product path fails loudly without a CUDA device instead of computing on the host."""
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    x = torch.rand(1, 3, 64, 64)
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        pnet.residual_model(x)
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        pnet(x, jpeg=(x, 0.1))
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        pnet.residual_model.compress(x)


def test_product_sources_never_import_the_oracle():
    import os
    from conftest import ROOT
    pkg = os.path.join(ROOT, "hyres-residual-enhanced-hybrid-image-compression_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "hyres_oracle" not in src and "rans_oracle" not in src and "from oracle" not in src, f


def test_jpeg_stage_contract(build_lib):
    """models/utils/turbo_jpeg_compression.py:17-77: .byte() truncation, bpp from buffer lengths."""
    import hyres_b200
    j = hyres_b200.TurboJPEGCompression(quality=1)
    x = (torch.arange(3 * 32 * 32).reshape(1, 3, 32, 32) % 256).float() / 255
    dec, bpp = j(x)
    assert dec.shape == x.shape and 0 <= dec.min() and dec.max() <= 1
    bufs = j.compress(x)
    assert bpp == pytest.approx(len(bufs[0].getvalue()) * 8 / (32 * 32))
    assert torch.equal(j.decompress(bufs, "cpu"), dec)


def test_container_roundtrip_and_errors(build_lib):
    """hyres_b200.container: pack / unpack of the compress() dict are exact inverses; malformed input raises."""
    import io
    import torch
    from hyres_b200 import container
    c = {"strings": [[[b"\x01\x02\x03\x04" * 3, b""], [b"abcdabcd", b"\x00" * 8]], [b"zzzzzzzz", b"yyyyyyyy"]],
         "shape": torch.Size([22, 16]), "time": 0.1, "jpeg_buffers": [io.BytesIO(b"\xff\xd8jpeg0"), io.BytesIO(b"\xff\xd8jpeg1")]}
    blob = container.pack(c)
    d = container.unpack(blob)
    assert d["strings"] == c["strings"] and d["shape"] == c["shape"]
    assert [b.getvalue() for b in d["jpeg_buffers"]] == [b.getvalue() for b in c["jpeg_buffers"]]
    assert container.pack(d) == blob
    no_jpeg = {k: v for k, v in c.items() if k != "jpeg_buffers"}
    d2 = container.unpack(container.pack(no_jpeg))
    assert "jpeg_buffers" not in d2 and d2["strings"] == c["strings"]
    for bad in (b"", b"XXXX" + blob[4:], blob[:-1], blob + b"\x00"):
        with pytest.raises(ValueError):
            container.unpack(bad)
    with pytest.raises(ValueError):
        container.pack({"strings": [[[b"a"], [b"b", b"c"]], [b"z"]], "shape": (1, 1)})
    # the header records which trunk arithmetic made the strings; a mismatching decoder is refused
    assert "trunk" not in d
    for tag, want in (("fp32x3", "fp32"), ("fp32h2", "fp32"), ("fp32x2", "fp32"), ("bf16", "bf16")):
        dt = container.unpack(container.pack(c, trunk=tag))
        assert dt["trunk"] == want and dt["strings"] == c["strings"]
        assert container.pack(dt) == container.pack(c, trunk=tag)  # the tag survives a second round trip
    container.check_trunk(container.unpack(container.pack(c, trunk="fp32x3")), "fp32x2")
    container.check_trunk(d, "bf16")  # untagged: accepted
    with pytest.raises(ValueError):
        container.check_trunk(container.unpack(container.pack(c, trunk="bf16")), "fp32x3")
    with pytest.raises(ValueError):
        container.pack(c, trunk="fp64")
