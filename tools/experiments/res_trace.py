"""Phase timeline of the resident-weights conv kernel (CTA 0): clock64 stamps per tile.
slots: 0 producer: load issued | MMA warp: 1 operands landed, 2 accumulator free, 3 MMAs issued |
       epilogue group (row 0): 4 loop top, 5 accumulator seen, 6 chunks done, store warp: 8 stage released
usage: python tools/experiments/res_trace.py [case]   (cases of tools/prof_kernels.py)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
buf = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
os.environ["HYRES_RES_TRACE"] = hex(buf.data_ptr())
import prof_kernels  # noqa: E402
for case in (sys.argv[1] if len(sys.argv) > 1 else "fus0lin").split(","):
    fn, _, _ = prof_kernels.make(case)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t = buf.cpu().view(64, 16).clone()
    buf.zero_()
    names = {0: "ld", 1: "m:A", 2: "m:accfree", 3: "m:iss", 4: "e:top", 5: "e:acc", 6: "e:chunks", 8: "s:rel"}
    print("==", case)
    have = [i for i in range(63) if int(t[i, 3]) and int(t[i + 1, 3])]
    for it in (have[20:30] if len(have) > 30 else have[:10]):
        base = int(t[it, 3])
        ev = sorted((int(t[it, k]) - base, names[k]) for k in names)
        print(f"tile {it} (MMA-issue period {int(t[it + 1, 3]) - base}): " + "  ".join(f"{n}@{c}" for c, n in ev))
