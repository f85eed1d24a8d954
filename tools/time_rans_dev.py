"""Device coder timing: ns per symbol of one warp's chain, alone and with many strings resident.
usage: python tools/time_rans_dev.py [n_symbols] [strings]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyres_b200 import coder, entropy, ops  # noqa: E402
from hyres_b200.models import get_scale_table  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1081344
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    RATE = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
    SIGMA = float(sys.argv[4]) if len(sys.argv) > 4 else 0.25
    gc = entropy.GaussianConditional(None)
    gc.update_scale_table(get_scale_table())
    t, dt = gc.tables(), gc.device_tables("cuda")
    rng = np.random.default_rng(0)
    # latents of a low-rate model: small scales, values near 0
    idx = np.minimum(rng.geometric(RATE, size=(S, n)) - 1, 63).astype(np.int32)
    last = t.sizes[idx] - 2
    sym = np.clip(np.rint(rng.normal(0, SIGMA, size=(S, n))), -(last // 2), last // 2).astype(np.int32)
    lay = coder.table_layout(t)
    value = sym - lay[1][idx]
    slots = (lay[0][idx] + value).astype(np.int32)
    known = np.broadcast_to(np.arange(n)[None, :] % 2 == 1, (S, n))
    codes = np.where(known, (1 << 30) | slots, idx).astype(np.int32)
    ds, dsl, dc = (torch.from_numpy(a).cuda() for a in (sym, slots, codes))
    t0 = time.perf_counter()
    want = coder.encode_batch(sym, slots, t, slots=True)
    host_ms = (time.perf_counter() - t0) * 1e3
    print(f"host encode: {host_ms:.1f} ms for {S} x {n} symbols ({sum(map(len, want)) * 8 / (S * n):.3f} bit/symbol)")
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        (got,) = ops.rans_encode_device([(ds, dsl, dt, True)])
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
    assert got == want
    print(f"device encode: {ms:.2f} ms -> {ms * 1e6 / n:.1f} ns per symbol of one string, {S} strings resident")
    words, table = ops.rans_upload([want], "cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    for name, c, flag in (("codes (every other symbol known)", dc, True), ("plain", torch.from_numpy(idx).cuda(), False)):
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dec = ops.rans_decode_device(words, table, 0, c, dt, flag, status)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
        assert int(status.item()) == 0
        ok = np.array_equal(dec.cpu().numpy()[~known], sym[~known]) if flag else np.array_equal(dec.cpu().numpy(), sym)
        print(f"device decode, {name}: {ms:.2f} ms -> {ms * 1e6 / n:.1f} ns per symbol, correct={ok}")
    t0 = time.perf_counter()
    coder.decode_batch(want, codes, t, codes=True, out=np.zeros_like(sym))
    print(f"host decode (codes): {(time.perf_counter() - t0) * 1e3:.1f} ms")


if __name__ == "__main__":
    main()
