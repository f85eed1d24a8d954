"""Host entropy coder of the product path (csrc/rans.cpp through the C-ABI).

Mirrors the pybind11 interface the reference crosses
(``compressai.ans.RansEncoder().encode_with_indexes`` / ``RansDecoder().decode_with_indexes``
and ``compressai._CXX.pmf_to_quantized_cdf``; reached from
models/checkerboard.py:159-165,172-173,206,261-267) but takes flat int32 numpy arrays
instead of Python lists, and adds batched calls that code independent strings in parallel.
"""
import ctypes as C
import os

import numpy as np

from . import _lib as L


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def pmf_to_quantized_cdf(pmf, precision=16):
    """-> int32 numpy array with len(pmf)+1 entries; raises ValueError on an invalid pmf."""
    p = np.ascontiguousarray(np.asarray(pmf, dtype=np.float32)).ravel()
    out = np.zeros(p.size + 1, dtype=np.uint32)
    rc = L.lib().hyres_pmf_to_quantized_cdf(p.ctypes.data, p.size, int(precision), out.ctypes.data)
    if rc != 0:
        raise ValueError("Invalid `pmf`: " + L.lib().hyres_last_error().decode())
    return out.astype(np.int32)


class CdfTables:
    """Flat view of an entropy model's ``_quantized_cdf`` / ``_cdf_length`` / ``_offset`` buffers."""

    def __init__(self, quantized_cdf, cdf_length, offset):
        self.cdf = _i32(quantized_cdf)
        if self.cdf.ndim != 2:
            raise ValueError(f"Invalid CDF size {self.cdf.shape}")
        self.sizes = _i32(cdf_length).ravel()
        self.offsets = _i32(offset).ravel()
        if self.sizes.size != self.cdf.shape[0] or self.offsets.size != self.cdf.shape[0]:
            raise ValueError("cdf_length / offset do not match the CDF table")


def encode_with_indexes(symbols, indexes, tables):
    s, ix = _i32(symbols).ravel(), _i32(indexes).ravel()
    if s.size != ix.size:
        raise ValueError("`symbols` and `indexes` should have the same size.")
    lib = L.lib()
    cap = int(lib.hyres_rans_encode_bound(s.size))
    out = np.empty(cap, dtype=np.uint8)
    n = C.c_int64(0)
    t = tables
    rc = lib.hyres_rans_encode(s.ctypes.data, ix.ctypes.data, s.size, t.cdf.ctypes.data, t.cdf.shape[0], t.cdf.shape[1],
                               t.sizes.ctypes.data, t.offsets.ctypes.data, out.ctypes.data, cap, C.byref(n))
    if rc != 0 and n.value > cap:
        cap = n.value
        out = np.empty(cap, dtype=np.uint8)
        rc = lib.hyres_rans_encode(s.ctypes.data, ix.ctypes.data, s.size, t.cdf.ctypes.data, t.cdf.shape[0],
                                   t.cdf.shape[1], t.sizes.ctypes.data, t.offsets.ctypes.data, out.ctypes.data, cap,
                                   C.byref(n))
    L.check(rc, "hyres_rans_encode")
    return out[: n.value].tobytes()


def decode_with_indexes(string, indexes, tables):
    ix = _i32(indexes).ravel()
    buf = np.frombuffer(string, dtype=np.uint8)
    out = np.empty(ix.size, dtype=np.int32)
    t = tables
    L.check(L.lib().hyres_rans_decode(buf.ctypes.data, buf.size, ix.ctypes.data, ix.size, t.cdf.ctypes.data,
                                      t.cdf.shape[0], t.cdf.shape[1], t.sizes.ctypes.data, t.offsets.ctypes.data,
                                      out.ctypes.data), "hyres_rans_decode")
    return out


def _threads(n):
    return max(1, min(n, os.cpu_count() or 1))


def table_layout(tables):
    """-> int32 [3, n_cdfs]: first packed entry, offset and escape bin of every CDF row (hyres_rans_table_layout): what
    the device-side front-end needs to emit coder slots / codes instead of (symbol, index) pairs."""
    t = tables
    out = np.zeros((3, t.cdf.shape[0]), dtype=np.int32)
    L.check(L.lib().hyres_rans_table_layout(t.cdf.ctypes.data, t.cdf.shape[0], t.cdf.shape[1], t.sizes.ctypes.data,
                                            t.offsets.ctypes.data, out.ctypes.data), "hyres_rans_table_layout")
    return out


def encode_batch(symbols, indexes, tables, threads=None, slots=False):
    """symbols / indexes: int32 arrays [count, n] (rows = independent strings), or equally long lists of such
    arrays (their rows are coded as one batch, without concatenating them) -> list of bytes.
    ``slots=True``: ``indexes`` holds coder slots (see ``table_layout`` / hyres_rans_encode_slots_batch); same bytes."""
    if isinstance(symbols, (list, tuple)):
        s_list, ix_list = [_i32(a) for a in symbols], [_i32(a) for a in indexes]
    else:
        s_list, ix_list = [_i32(symbols)], [_i32(indexes)]
    if len(s_list) != len(ix_list):
        raise ValueError("`symbols` and `indexes` should have the same size.")
    rows_s, rows_i = [], []
    for s, ix in zip(s_list, ix_list):
        if s.shape != ix.shape or s.ndim != 2:
            raise ValueError("`symbols` and `indexes` should be [count, n] arrays of the same size.")
        rows_s += [s[i] for i in range(s.shape[0])]
        rows_i += [ix[i] for i in range(s.shape[0])]
    count = len(rows_s)
    if count == 0:
        return []
    lib = L.lib()
    ns = np.array([r.size for r in rows_s], dtype=np.int64)
    caps = np.array([int(lib.hyres_rans_encode_bound(int(n))) for n in ns], dtype=np.int64)
    t = tables
    while True:
        outs = [np.empty(int(c), dtype=np.uint8) for c in caps]
        sp = (C.c_void_p * count)(*[r.ctypes.data for r in rows_s])
        ip = (C.c_void_p * count)(*[r.ctypes.data for r in rows_i])
        op = (C.c_void_p * count)(*[o.ctypes.data for o in outs])
        lens = np.zeros(count, dtype=np.int64)
        fn = lib.hyres_rans_encode_slots_batch if slots else lib.hyres_rans_encode_batch
        rc = fn(count, sp, ip, ns.ctypes.data, t.cdf.ctypes.data, t.cdf.shape[0],
                t.cdf.shape[1], t.sizes.ctypes.data, t.offsets.ctypes.data, op,
                caps.ctypes.data, lens.ctypes.data, _threads(threads or count))
        if rc != 0 and (lens > caps).any():
            caps = np.maximum(caps, lens)
            continue
        L.check(rc, "hyres_rans_encode_batch")
        return [outs[i][: lens[i]].tobytes() for i in range(count)]


def decode_batch(strings, indexes, tables, threads=None, out=None, codes=False):
    """strings: list of bytes; indexes int32 [count, n] -> int32 [count, n] (``out``: preallocated result).
    ``codes=True``: ``indexes`` holds decoder codes (hyres_rans_decode_codes_batch): entries of known symbols only
    advance the coder and leave ``out`` untouched there."""
    ix = _i32(indexes)
    count, n = ix.shape
    if len(strings) != count:
        raise ValueError("Invalid strings or indexes parameters")
    if count == 0:
        return np.empty((0, n), dtype=np.int32)
    bufs = [np.frombuffer(s, dtype=np.uint8) for s in strings]
    if out is None:
        out = np.empty((count, n), dtype=np.int32)
    elif out.shape != (count, n) or out.dtype != np.int32 or not out.flags["C_CONTIGUOUS"]:
        raise ValueError("decode_batch: `out` must be a C-contiguous int32 [count, n] array")
    bp = (C.c_void_p * count)(*[b.ctypes.data for b in bufs])
    ip = (C.c_void_p * count)(*[ix[i].ctypes.data for i in range(count)])
    op = (C.c_void_p * count)(*[out[i].ctypes.data for i in range(count)])
    lens = np.array([b.size for b in bufs], dtype=np.int64)
    ns = np.full(count, n, dtype=np.int64)
    t = tables
    fn = L.lib().hyres_rans_decode_codes_batch if codes else L.lib().hyres_rans_decode_batch
    L.check(fn(count, bp, lens.ctypes.data, ip, ns.ctypes.data, t.cdf.ctypes.data, t.cdf.shape[0], t.cdf.shape[1],
               t.sizes.ctypes.data, t.offsets.ctypes.data, op, _threads(threads or count)), "hyres_rans_decode_batch")
    return out


def host_cores_per_process():
    """Host cores this process can count on: all of them divided by the processes torchrun started on this node (one
    per GPU); ``HYRES_HOST_CORES`` overrides (the same rule as the coder's thread pool, csrc/rans.cpp)."""
    env = os.environ.get("HYRES_HOST_CORES")
    if env:
        return max(1, int(env))
    hw = os.cpu_count() or 1
    return max(1, hw // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)))


class DeviceTables:
    """The coder's packed tables on a CUDA device (hyres_rans_table_export), for the device-resident coder
    (csrc/rans_dev.cu): ``enc`` 16-byte encoder entries, ``sf`` decoder words, ``rows`` int32 [4, n_cdfs]."""

    def __init__(self, tables, device):
        import torch
        t = tables
        lib = L.lib()
        args = (t.cdf.ctypes.data, t.cdf.shape[0], t.cdf.shape[1], t.sizes.ctypes.data, t.offsets.ctypes.data)
        n = int(lib.hyres_rans_table_entries(*args))
        if n <= 0:
            raise L.HyresError("hyres_rans_table_entries failed: " + lib.hyres_last_error().decode())
        enc = np.zeros(n * 16, dtype=np.uint8)
        sf = np.zeros(n, dtype=np.uint32)
        rows = np.zeros((4, t.cdf.shape[0]), dtype=np.int32)
        L.check(lib.hyres_rans_table_export(*args, enc.ctypes.data, sf.ctypes.data, rows.ctypes.data),
                "hyres_rans_table_export")
        self.n_entries, self.n_rows = n, int(t.cdf.shape[0])
        self.enc = torch.from_numpy(enc).to(device)
        self.sf = torch.from_numpy(sf.view(np.int32)).to(device)
        self.rows = torch.from_numpy(rows).to(device)
