"""Phase timeline of the fused ResidualUnit kernel (CTA 0): clock64 stamps per tile.
slots: E-warp 0: 0 loop top, 1 G1_DONE seen, 2 E1 done, 3 G2_DONE seen, 4 E2 done, 5 G3_DONE seen, 6 E3 staged
       MMA lane: 8 T1_READY seen, 9 G2 issued, 10 T2_READY(+ACC_FREE) seen, 11 G3 issued"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
buf = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
os.environ["HYRES_RU_TRACE"] = hex(buf.data_ptr())
from hyres_b200 import ops
g = torch.Generator().manual_seed(1)
c1 = ops.ConvLayer(torch.randn(64, 128, 1, 1, generator=g) * 0.1, torch.zeros(64))
c2 = ops.ConvLayer(torch.randn(64, 64, 3, 3, generator=g) * 0.05, torch.zeros(64), pad=1)
c3 = ops.ConvLayer(torch.randn(128, 64, 1, 1, generator=g) * 0.1, torch.zeros(128))
x = torch.randn(16, 256, 384, 128, device="cuda").bfloat16()
out = torch.empty_like(x)
for _ in range(3):
    ops.ru_fused(x, c1, c2, c3, True, out=out)
torch.cuda.synchronize()
t = buf.cpu().view(64, 16)
names = {0: "top", 1: "G1seen", 2: "E1done", 3: "G2seen", 4: "E2done", 5: "G3seen", 6: "E3done", 8: "m:T1seen", 9: "m:G2iss", 10: "m:T2seen", 11: "m:G3iss"}
for it in range(20, 28):
    base = int(t[it, 0])
    ev = sorted((int(t[it, k]) - base, names[k]) for k in names)
    print(f"tile {it} (period {int(t[it + 1, 0]) - base}): " + "  ".join(f"{n}@{c}" for c, n in ev))
