"""Entropy coder: the product's C++ coder (csrc/rans.cpp, host code behind the C-ABI) against the
plain-C oracle and the committed known-answer vector; round-trip properties; error paths.
CPU only (the coder is host code on both sides)."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from conftest import load_golden


@pytest.fixture(scope="module")
def tables(build_lib, oracle_net):
    from hyres_b200 import coder
    gc = oracle_net.residual_model.gaussian_conditional
    return coder.CdfTables(gc._quantized_cdf.numpy(), gc._cdf_length.numpy(), gc._offset.numpy())


def test_known_answer_vector(build_lib, oracle, tables):
    from hyres_b200 import coder
    g = load_golden("rans_kat")
    sym, idx, want = g["symbols"].astype(np.int32), g["indexes"].astype(np.int32), g["string"].tobytes()
    # tables regenerated here must be the fixture's
    assert (tables.sizes == g["cdf_length"]).all() and (tables.offsets == g["offset"]).all()
    assert (tables.cdf[0, :16] == g["cdf_row0"]).all() and (tables.cdf[63, :8] == g["cdf_row63_head"]).all()
    assert oracle.rans_encode_with_indexes(sym, idx, tables.cdf, tables.sizes, tables.offsets) == want
    mine = coder.encode_with_indexes(sym, idx, tables)
    assert mine == want
    assert len(mine) % 4 == 0 and len(mine) >= 8
    assert (coder.decode_with_indexes(want, idx, tables) == sym).all()
    assert (oracle.rans_decode_with_indexes(mine, idx, tables.cdf, tables.sizes, tables.offsets) == sym).all()


def test_empty_input_is_the_flushed_state(build_lib, oracle, tables):
    from hyres_b200 import coder
    e = np.zeros(0, dtype=np.int32)
    mine = coder.encode_with_indexes(e, e, tables)
    assert mine == oracle.rans_encode_with_indexes(e, e, tables.cdf, tables.sizes, tables.offsets)
    assert mine == (1 << 31).to_bytes(8, "little")
    assert coder.decode_with_indexes(mine, e, tables).size == 0


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 700), st.floats(0.2, 40.0), st.booleans())
def test_roundtrip_and_byte_identity(build_lib, oracle, tables, seed, n, spread, escapes):
    from hyres_b200 import coder
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, 64, size=n).astype(np.int32)
    sym = np.round(rng.standard_normal(n) * spread).astype(np.int32)
    if escapes:
        k = rng.integers(0, n, size=max(1, n // 9))
        sym[k] += rng.integers(-70000, 70000, size=k.size).astype(np.int32)
    want = oracle.rans_encode_with_indexes(sym, idx, tables.cdf, tables.sizes, tables.offsets)
    mine = coder.encode_with_indexes(sym, idx, tables)
    assert mine == want
    assert (coder.decode_with_indexes(mine, idx, tables) == sym).all()


def test_batch_matches_single(build_lib, tables):
    from hyres_b200 import coder
    rng = np.random.default_rng(7)
    count, n = 5, 3000
    idx = rng.integers(0, 64, size=(count, n)).astype(np.int32)
    sym = np.round(rng.standard_normal((count, n)) * 3).astype(np.int32)
    sym[2, ::50] = 9999
    strings = coder.encode_batch(sym, idx, tables, threads=3)
    assert strings == [coder.encode_with_indexes(sym[i], idx[i], tables) for i in range(count)]
    assert (coder.decode_batch(strings, idx, tables, threads=2) == sym).all()
    assert coder.encode_batch(sym[:0], idx[:0], tables) == []


@pytest.mark.parametrize("group", [2, 4])
def test_lock_step_groups_code_the_same_bytes(build_lib, group):
    """Two / four strings coded in lock step by one thread (csrc/rans.cpp: encode_n / decode_n; chosen when the host
    cores are oversubscribed, forced here through HYRES_RANS_GROUP) give the bytes and symbols of single calls,
    escapes and unequal lengths included."""
    import os
    import subprocess
    import sys
    code = r"""
import sys
import numpy as np
sys.path.insert(0, %r)
import hyres_b200
from hyres_b200 import coder
gc = hyres_b200.models.GaussianConditional(None)
gc.update_scale_table(hyres_b200.get_scale_table())
t = gc.tables()
rng = np.random.default_rng(11)
count, n = 7, 20000
idx = rng.integers(0, 64, size=(count, n)).astype(np.int32)
sym = np.round(rng.standard_normal((count, n)) * rng.choice([0.3, 2.0, 30.0], size=(count, 1))).astype(np.int32)
sym[1, ::37] += 70000
sym[4, ::91] -= 70000
single = [coder.encode_with_indexes(sym[i], idx[i], t) for i in range(count)]
assert coder.encode_batch(sym, idx, t, threads=2) == single
assert (coder.decode_batch(single, idx, t, threads=2) == sym).all()
# unequal lengths fall back to smaller groups
rows_s = [sym[0], sym[1][:1000], sym[2], sym[3][:1000], sym[5]]
rows_i = [idx[0], idx[1][:1000], idx[2], idx[3][:1000], idx[5]]
got = coder.encode_batch([r[None] for r in rows_s], [r[None] for r in rows_i], t)
assert got == [coder.encode_with_indexes(a, b, t) for a, b in zip(rows_s, rows_i)]
print("ok")
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, HYRES_RANS_GROUP=str(group))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


def test_pmf_to_quantized_cdf_matches_oracle(build_lib, oracle):
    from hyres_b200 import coder
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 17, 300, 3133):
        p = rng.random(n).astype(np.float32) ** 6  # many tiny bins -> exercises the steal loop
        p /= p.sum()
        got = coder.pmf_to_quantized_cdf(p)
        want = oracle.pmf_to_quantized_cdf(p).numpy()
        assert (got == want).all()
        assert got[0] == 0 and got[-1] == 65536 and (np.diff(got) > 0).all()
    with pytest.raises(ValueError):
        coder.pmf_to_quantized_cdf(np.array([0.5, -0.1, 0.6], dtype=np.float32))
    with pytest.raises(ValueError):
        coder.pmf_to_quantized_cdf(np.array([0.5, np.nan], dtype=np.float32))


def test_error_paths(build_lib, tables):
    from hyres_b200 import _lib, coder
    sym = np.zeros(8, dtype=np.int32)
    bad_idx = np.full(8, 64, dtype=np.int32)  # out-of-range CDF row
    with pytest.raises(_lib.HyresError):
        coder.encode_with_indexes(sym, bad_idx, tables)
    with pytest.raises(ValueError):
        coder.encode_with_indexes(sym, np.zeros(7, dtype=np.int32), tables)
    with pytest.raises(_lib.HyresError):  # a stream shorter than the 8-byte flushed state
        coder.decode_with_indexes(b"\x00\x01\x02", np.zeros(4, dtype=np.int32), tables)
    # output capacity too small: HYRES_ERR_ARG and the needed length reported
    lib = _lib.lib()
    s = np.arange(100, dtype=np.int32) % 5
    ix = np.zeros(100, dtype=np.int32)
    out = np.empty(4, dtype=np.uint8)
    n = C.c_int64(0)
    rc = lib.hyres_rans_encode(s.ctypes.data, ix.ctypes.data, s.size, tables.cdf.ctypes.data, tables.cdf.shape[0],
                               tables.cdf.shape[1], tables.sizes.ctypes.data, tables.offsets.ctypes.data,
                               out.ctypes.data, 4, C.byref(n))
    assert rc == -1 and n.value > 4


def test_entropy_model_string_api(build_lib, oracle_net):
    """EntropyModel.encode_symbols / decode_symbols keep compressai's argument checks."""
    import torch
    import hyres_b200
    gc = hyres_b200.models.GaussianConditional(None)
    with pytest.raises(ValueError, match="Uninitialized CDFs"):
        gc.encode_symbols(torch.zeros(1, 4, dtype=torch.int32), torch.zeros(1, 4, dtype=torch.int32))
    gc.update_scale_table(hyres_b200.get_scale_table())
    ogc = oracle_net.residual_model.gaussian_conditional
    assert torch.equal(gc._quantized_cdf, ogc._quantized_cdf)
    assert torch.equal(gc._cdf_length, ogc._cdf_length) and torch.equal(gc._offset, ogc._offset)
    with pytest.raises(ValueError):
        gc.encode_symbols(torch.zeros(4, dtype=torch.int32), torch.zeros(4, dtype=torch.int32))
    with pytest.raises(ValueError):
        gc.encode_symbols(torch.zeros(1, 4, dtype=torch.int32), torch.zeros(1, 5, dtype=torch.int32))
    sym = torch.randint(-5, 6, (2, 3, 4, 4), dtype=torch.int32)
    idx = torch.randint(0, 64, (2, 3, 4, 4), dtype=torch.int32)
    strings = gc.encode_symbols(sym, idx)
    assert len(strings) == 2 and all(isinstance(s, bytes) for s in strings)
    assert torch.equal(gc.decode_symbols(strings, idx), sym)
    with pytest.raises(ValueError):
        gc.decode_symbols(strings[:1], idx)


def test_slots_and_codes_front_end(build_lib, tables):
    """The device front-end's streams (csrc/rans.cpp: slots for the encoder, codes with known symbols for the decoder),
    built here with numpy from (symbol, index): the same bytes as the plain encoder, and a decoder that skips the search
    for known symbols stays in step and returns every unknown symbol."""
    from hyres_b200 import coder
    rng = np.random.default_rng(21)
    count, n = 4, 6000
    idx = rng.integers(0, 64, size=(count, n)).astype(np.int32)
    sym = np.round(rng.standard_normal((count, n)) * rng.choice([0.4, 3.0, 40.0], size=(count, 1))).astype(np.int32)
    sym[0, ::41] += 70000  # escapes
    sym[2, ::53] -= 70000
    base, off, last = coder.table_layout(tables)
    assert (off == tables.offsets).all() and (last == tables.sizes - 2).all() and (np.diff(base) > 0).all()
    value = sym - off[idx]
    inside = (value >= 0) & (value < last[idx])
    slots = np.where(inside, base[idx] + value, -(idx + 1)).astype(np.int32)
    plain = coder.encode_batch(sym, idx, tables)
    assert coder.encode_batch(sym, slots, tables, slots=True) == plain
    # decoder: every second position is "known" (when its value is inside the table)
    known = inside & (np.arange(n)[None, :] % 2 == 0)
    codes = np.where(known, (1 << 30) | (base[idx] + value), idx).astype(np.int32)
    out = np.full((count, n), -12345, dtype=np.int32)
    coder.decode_batch(plain, codes, tables, out=out, codes=True)
    assert (out[~known] == sym[~known]).all() and (out[known] == -12345).all()
    # malformed slot / code
    from hyres_b200 import _lib
    bad = slots.copy()
    bad[1, 5] = 2 ** 29
    with pytest.raises(_lib.HyresError):
        coder.encode_batch(sym, bad, tables, slots=True)


def test_device_table_export_is_the_host_coders_tables(build_lib, tables):
    """hyres_rans_table_export (what the device-resident coder uploads): rows agree with hyres_rans_table_layout, the
    decoder words are (start, freq - 1) of every bin, and the encoder entries divide exactly: for random states x,
    x + bias + (mulhi(x, rcp) >> shift) * (2^16 - freq) == ((x // freq) << 16) + x % freq + start."""
    from hyres_b200 import _lib, coder
    t = tables
    lib = _lib.lib()
    args = (t.cdf.ctypes.data, t.cdf.shape[0], t.cdf.shape[1], t.sizes.ctypes.data, t.offsets.ctypes.data)
    n = int(lib.hyres_rans_table_entries(*args))
    assert n == int(np.maximum(1, np.minimum(t.sizes, t.cdf.shape[1])).sum())
    enc = np.zeros(n * 16, dtype=np.uint8)
    sf = np.zeros(n, dtype=np.uint32)
    rows = np.zeros((4, t.cdf.shape[0]), dtype=np.int32)
    assert lib.hyres_rans_table_export(*args, enc.ctypes.data, sf.ctypes.data, rows.ctypes.data) == 0
    lay = coder.table_layout(t)
    assert (rows[3] == 1).all() and (rows[:3] == lay).all()
    e = enc.view(np.dtype([("rcp", "<u8"), ("bias", "<u4"), ("freq_m1", "<u2"), ("shift", "u1"), ("valid", "u1")]))
    rng = np.random.default_rng(0)
    for r in (0, 7, 31, 63):
        base, last = int(rows[0, r]), int(rows[2, r])
        cdf = t.cdf[r].astype(np.int64)
        for v in sorted({0, 1, last // 2, last - 1, last}):
            start, freq = int(cdf[v]), int(cdf[v + 1] - cdf[v])
            assert int(sf[base + v]) == start | ((freq - 1) << 16)
            ent = e[base + v]
            assert ent["valid"] == 1 and int(ent["freq_m1"]) + 1 == freq
            for x in [1 << 31, (freq << 47) - 1] + [int(rng.integers(1 << 31, freq << 47)) for _ in range(50)]:
                q = ((x * int(ent["rcp"])) >> 64) >> int(ent["shift"])
                got = x + int(ent["bias"]) + q * (65536 - freq)
                assert got == ((x // freq) << 16) + x % freq + start, (r, v, x)
    assert lib.hyres_rans_table_export(*args, None, sf.ctypes.data, rows.ctypes.data) != 0


def test_coder_choice_follows_the_host_cores_of_the_process(build_lib, monkeypatch):
    """coder = "auto": the device coder when this process can count on fewer than 16 host cores (one process per GPU
    under torchrun share the box), the host coder otherwise; explicit choices are respected; nonsense is refused."""
    import hyres_b200
    from hyres_b200 import coder
    net = hyres_b200.LightWeightCheckerboard()
    monkeypatch.setenv("HYRES_HOST_CORES", "32")
    assert coder.host_cores_per_process() == 32 and not net.uses_device_coder()
    monkeypatch.setenv("HYRES_HOST_CORES", "4")
    assert net.uses_device_coder()
    monkeypatch.delenv("HYRES_HOST_CORES")
    monkeypatch.setattr("os.cpu_count", lambda: 32)
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "8")
    assert coder.host_cores_per_process() == 4 and net.uses_device_coder()
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "1")
    assert not net.uses_device_coder()
    net.coder = "device"
    assert net.uses_device_coder()
    net.coder = "host"
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "8")
    assert not net.uses_device_coder()
    net.coder = "gpu"
    with pytest.raises(ValueError):
        net.uses_device_coder()
