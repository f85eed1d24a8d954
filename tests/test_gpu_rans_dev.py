"""The device-resident entropy coder (csrc/rans_dev.cu): one warp per string, the same bytes as the host coder.

Bar: byte-identical strings against ``csrc/rans.cpp`` (which is itself held to the plain-C oracle and the known-answer
vector in tests/test_rans.py) for plain (symbol, CDF row) input and for the device front-end's slots, with escapes,
at ragged lengths; the decoder returns the coded symbols from the host coder's strings and from its own, in plain and
in codes mode (known symbols only move the state); through the model, ``compress`` / ``decompress`` with
``coder = "device"`` give the strings and the reconstruction of ``coder = "host"`` bit for bit."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gc_tables(build_lib):
    from hyres_b200 import entropy
    from hyres_b200.models import get_scale_table
    gc = entropy.GaussianConditional(None)
    gc.update_scale_table(get_scale_table())
    return gc


def _symbols(tables, rows, n, rng, escape_rate):
    """Random symbols: mostly inside each row's table (concentrated near 0 like real latents, with a wide tail),
    a fraction ``escape_rate`` outside it on either side."""
    idx = rng.integers(0, rows, size=n).astype(np.int32)
    last = tables.sizes[idx] - 2
    spread = np.maximum(1, (last // 2) * rng.choice([0.02, 0.1, 1.0], size=n, p=[0.6, 0.3, 0.1]))
    val = np.clip(np.rint(rng.normal(0, 1, size=n) * spread / 2), -(last // 2), last - last // 2 - 1).astype(np.int64)
    sym = val  # offsets are -(last // 2) for the Gaussian tables: value 0 sits at the row's centre
    esc = rng.random(n) < escape_rate
    far = rng.integers(1, 70000, size=n) * rng.choice([-1, 1], size=n)
    sym = np.where(esc, np.where(far > 0, last - last // 2 + far, -(last // 2) - 1 + far), sym)
    return sym.astype(np.int32), idx


@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000, 70001])
@pytest.mark.parametrize("escape_rate", [0.0, 0.02])
def test_device_coder_matches_host_bytes(gc_tables, n, escape_rate):
    from hyres_b200 import coder, ops
    t = gc_tables.tables()
    dt = gc_tables.device_tables("cuda")
    rng = np.random.default_rng(n * 7 + int(escape_rate * 100))
    B = 5
    syms, idxs = zip(*[_symbols(t, t.cdf.shape[0], n, rng, escape_rate) for _ in range(B)])
    sym, idx = np.stack(syms), np.stack(idxs)
    want = coder.encode_batch(sym, idx, t)
    ds, di = torch.from_numpy(sym).cuda(), torch.from_numpy(idx).cuda()
    (got,) = ops.rans_encode_device([(ds, di, dt, False)])
    assert got == want
    # slots: the packed entry of an in-table value, -(row + 1) otherwise (what hyres_gc_symbols emits)
    lay = coder.table_layout(t)
    value = sym - lay[1][idx]
    inside = (value >= 0) & (value < lay[2][idx])
    slots = np.where(inside, lay[0][idx] + value, -(idx + 1)).astype(np.int32)
    assert coder.encode_batch(sym, slots, t, slots=True) == want
    (got_slots,) = ops.rans_encode_device([(ds, torch.from_numpy(slots).cuda(), dt, True)])
    assert got_slots == want
    # decoding: the host coder's strings and (the same bytes) its own
    words, table = ops.rans_upload([want], "cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    dec = ops.rans_decode_device(words, table, 0, di, dt, False, status)
    assert int(status.item()) == 0
    assert np.array_equal(dec.cpu().numpy(), sym)
    # codes: every other symbol is known to the caller (bit 30 | packed entry of its bin, escapes stay unknown)
    known = inside & (np.arange(n)[None, :] % 2 == 1)
    codes = np.where(known, (1 << 30) | (lay[0][idx] + value), idx).astype(np.int32)
    host = coder.decode_batch(want, codes, t, codes=True, out=np.full(sym.shape, -12345, dtype=np.int32))
    dec = ops.rans_decode_device(words, table, 0, torch.from_numpy(codes).cuda(), dt, True, status).cpu().numpy()
    assert int(status.item()) == 0
    assert np.array_equal(dec[~known], sym[~known]) and np.array_equal(host[~known], sym[~known])


def test_device_coder_wide_rows_take_the_search_path(gc_tables):
    """Symbols spread over the widest rows (3133 bins): most fall outside the 32-bin window around the centre."""
    from hyres_b200 import coder, ops
    t = gc_tables.tables()
    dt = gc_tables.device_tables("cuda")
    rng = np.random.default_rng(5)
    n, B = 20000, 3
    idx = rng.integers(56, 64, size=(B, n)).astype(np.int32)
    last = t.sizes[idx] - 2
    sym = (rng.integers(0, 1 << 30, size=(B, n)) % last - last // 2).astype(np.int32)
    want = coder.encode_batch(sym, idx, t)
    (got,) = ops.rans_encode_device([(torch.from_numpy(sym).cuda(), torch.from_numpy(idx).cuda(), dt, False)])
    assert got == want
    words, table = ops.rans_upload([want], "cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    dec = ops.rans_decode_device(words, table, 0, torch.from_numpy(idx).cuda(), dt, False, status)
    assert int(status.item()) == 0 and np.array_equal(dec.cpu().numpy(), sym)


def test_device_coder_streams_made_of_escapes(gc_tables):
    """Every symbol outside its table: the stream outgrows the encoder's first working-space estimate (half a word per
    symbol), the wrapper retries with room for escapes, and the bytes are still the host coder's."""
    from hyres_b200 import coder, ops
    t, dt = gc_tables.tables(), gc_tables.device_tables("cuda")
    rng = np.random.default_rng(13)
    B, n = 3, 20000
    idx = rng.integers(0, 64, size=(B, n)).astype(np.int32)
    last = t.sizes[idx] - 2
    far = rng.integers(1, 1 << 20, size=(B, n))
    sym = np.where(rng.random((B, n)) < 0.5, last - last // 2 + far, -(last // 2) - 1 - far).astype(np.int32)
    want = coder.encode_batch(sym, idx, t)
    assert min(len(w) for w in want) > 4 * (n // 2 + 4096)
    (got,) = ops.rans_encode_device([(torch.from_numpy(sym).cuda(), torch.from_numpy(idx).cuda(), dt, False)])
    assert got == want
    words, table = ops.rans_upload([want], "cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    dec = ops.rans_decode_device(words, table, 0, torch.from_numpy(idx).cuda(), dt, False, status)
    assert int(status.item()) == 0 and np.array_equal(dec.cpu().numpy(), sym)


def test_device_coder_many_strings_and_groups(gc_tables):
    """More strings than one block holds (8 warps per block) and more groups than one launch takes (4): strings of
    different lengths and table sets side by side, every one byte-identical to the host coder's."""
    from hyres_b200 import coder, entropy, ops
    torch.manual_seed(5)
    eb = entropy.EntropyBottleneck(16)
    eb.update(force=True)
    t, dt = gc_tables.tables(), gc_tables.device_tables("cuda")
    tz, dtz = eb.tables(), eb.device_tables("cuda")
    rng = np.random.default_rng(9)
    groups, want = [], []
    for g, (B, n) in enumerate([(20, 333), (1, 70), (9, 1024), (3, 5), (11, 640), (2, 4097)]):
        if g % 3 == 1:
            idx = rng.integers(0, 16, size=(B, n)).astype(np.int32)
            sym = rng.integers(-30, 31, size=(B, n)).astype(np.int32)
            tab, dtab = tz, dtz
        else:
            syms, idxs = zip(*[_symbols(t, t.cdf.shape[0], n, rng, 0.01) for _ in range(B)])
            sym, idx = np.stack(syms), np.stack(idxs)
            tab, dtab = t, dt
        want.append(coder.encode_batch(sym, idx, tab))
        groups.append((torch.from_numpy(sym).cuda(), torch.from_numpy(idx).cuda(), dtab, False))
    got = ops.rans_encode_device(groups)
    assert got == want
    words, table = ops.rans_upload(want, "cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    first = 0
    for (sym, idx, dtab, _), strings in zip(groups, want):
        dec = ops.rans_decode_device(words, table, first, idx, dtab, False, status)
        assert torch.equal(dec, sym)
        first += len(strings)
    assert int(status.item()) == 0


def test_device_coder_entropy_bottleneck_tables(build_lib):
    """The factorised prior's tables (one row per channel, short rows, non-symmetric offsets) and mixed groups."""
    from hyres_b200 import coder, entropy, ops
    torch.manual_seed(3)
    eb = entropy.EntropyBottleneck(128)
    eb.update(force=True)
    t, dt = eb.tables(), eb.device_tables("cuda")
    B, h, w = 3, 6, 5
    idx = eb._build_indexes((B, 128, h, w))
    rng = np.random.default_rng(11)
    sym = rng.integers(-25, 26, size=tuple(idx.shape)).astype(np.int32)
    want = coder.encode_batch(sym.reshape(B, -1), idx.numpy().reshape(B, -1), t)
    didx = eb.device_indexes((B, 128, h, w), "cuda")
    assert torch.equal(didx.cpu(), idx)
    (got,) = ops.rans_encode_device([(torch.from_numpy(sym).cuda(), didx, dt, False)])
    assert got == want
    words, table = ops.rans_upload([want], "cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    dec = ops.rans_decode_device(words, table, 0, didx, dt, False, status)
    assert int(status.item()) == 0 and np.array_equal(dec.cpu().numpy(), sym)


def test_device_decoder_flags_malformed_streams(gc_tables):
    from hyres_b200 import ops
    dt = gc_tables.device_tables("cuda")
    with pytest.raises(ValueError):
        ops.rans_upload([[b"\x00" * 7]], "cuda")
    # an escape whose nibble count says "more than 32 raw bits": every row's last bin is the escape
    t = gc_tables.tables()
    last = int(t.sizes[0]) - 2
    start = int(t.cdf[0, last])
    state = (1 << 40) | start  # cum = start of the escape bin; the next nibbles come from the high bits
    bad = (state | (0xFFFF << 16)).to_bytes(8, "little") + b"\xff" * 64
    words, table = ops.rans_upload([[bad]], "cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.rans_decode_device(words, table, 0, torch.zeros((1, 4), dtype=torch.int32, device="cuda"), dt, False, status)
    assert int(status.item()) != 0
    # a row index outside the tables
    status.zero_()
    ops.rans_decode_device(words, table, 0, torch.full((1, 4), 64, dtype=torch.int32, device="cuda"), dt, False, status)
    assert int(status.item()) != 0


@pytest.fixture(scope="module")
def pnet(build_lib, oracle_net):
    import hyres_b200
    net = hyres_b200.ResidualJPEGCompression()
    net.load_state_dict(oracle_net.state_dict())
    return net.cuda().eval()


def test_model_strings_and_reconstruction_do_not_depend_on_the_coder(pnet, oracle):
    """compress / decompress with the strings coded on the device = with the host coder, bit for bit, in every
    combination (device-coded strings decode on the host and vice versa); through the JPEG wrapper too."""
    codec = pnet.residual_model
    x = oracle.synthetic_residual(3, 96, 160, seed=21).cuda()
    out = {}
    try:
        for c in ("host", "device"):
            codec.coder = c
            with torch.no_grad():
                out[c] = codec.compress(x)
        assert out["host"]["strings"] == out["device"]["strings"]
        assert out["host"]["shape"] == out["device"]["shape"]
        rec = {}
        for c in ("host", "device"):
            codec.coder = c
            with torch.no_grad():
                rec[c] = codec.decompress(out["host"]["strings"], out["host"]["shape"])["x_hat"]
        assert torch.equal(rec["host"], rec["device"])
        img = oracle.synthetic_image(2, 64, 96, seed=4).cuda()
        full = {}
        for c in ("host", "device"):
            codec.coder = c
            with torch.no_grad():
                cc = pnet.compress(img)
                full[c] = (cc["strings"], pnet.decompress(cc)["x_hat"])
        assert full["host"][0] == full["device"][0]
        assert torch.equal(full["host"][1], full["device"][1])
        codec.coder = "device"
        broken = [[out["host"]["strings"][0][0][:2], out["host"]["strings"][0][1]], out["host"]["strings"][1]]
        with pytest.raises(ValueError):
            codec.decompress(broken, out["host"]["shape"])
    finally:
        codec.coder = "auto"


@pytest.mark.parametrize("shape", [(1, 32, 32), (2, 64, 32), (1, 32, 96)])
def test_smallest_shapes_through_the_device_coder(pnet, oracle, shape):
    """The smallest legal inputs (H, W multiples of 32: one hyper-latent position per channel, strings shorter than a
    32-symbol chunk for z): device-coded strings and reconstruction equal the host coder's."""
    codec = pnet.residual_model
    B, H, W = shape
    x = oracle.synthetic_residual(B, H, W, seed=H + W).cuda()
    try:
        res = {}
        for c in ("host", "device"):
            codec.coder = c
            with torch.no_grad():
                cc = codec.compress(x)
                res[c] = (cc["strings"], cc["shape"], codec.decompress(cc["strings"], cc["shape"])["x_hat"])
        assert res["host"][0] == res["device"][0] and res["host"][1] == res["device"][1]
        assert torch.equal(res["host"][2], res["device"][2])
    finally:
        codec.coder = "auto"


def test_codec_pipeline_with_the_device_coder(pnet, oracle):
    """Several images in flight, CUDA-graph phases and coder warps on every worker's stream: same strings and pixels
    as single host-coder calls."""
    from hyres_b200.codec_pipeline import CodecPipeline
    codec = pnet.residual_model
    xs = [torch.rand(2, 3, 64, 96, generator=torch.Generator().manual_seed(s)).pin_memory() for s in range(6)]
    try:
        codec.coder = "host"
        with torch.no_grad():
            want = []
            for x in xs:
                c = pnet.compress(x.cuda())
                want.append((c["strings"], pnet.decompress(c)["x_hat"].cpu()))
        codec.coder = "device"
        pipe = CodecPipeline(pnet, workers=3)
        try:
            for rep in range(3):  # eager pass, graph capture, replay
                got = list(pipe.roundtrip(xs))
                for (c, x_hat), (strings, ref) in zip(got, want):
                    assert c["strings"] == strings
                    assert torch.equal(x_hat.cpu(), ref)
        finally:
            pipe.close()
    finally:
        codec.coder = "auto"
