"""Per-launch table of the conv kernel over one forward step (CUDA events around each launch).
usage: python tools/profile_layers.py [--B 16 --H 512 --W 768] -> gpurun_out/layers.json + stdout"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import hyres_b200  # noqa: E402
from hyres_b200 import ops, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=16)
    ap.add_argument("--H", type=int, default=512)
    ap.add_argument("--W", type=int, default=768)
    ap.add_argument("--out", default="gpurun_out/layers.json")
    a = ap.parse_args()
    torch.manual_seed(1926)
    net = hyres_b200.ResidualJPEGCompression()
    net.update(force=True)
    net = net.cuda().eval()
    x = synthetic.synthetic_image(a.B, a.H, a.W).cuda()
    jd = (x * 0.9 + 0.05)
    with torch.no_grad():
        for _ in range(2):
            net(x, jpeg=(jd, 0.3))
        torch.cuda.synchronize()
        ops.ConvLayer.profile_begin()
        net(x, jpeg=(jd, 0.3))
        tot, n = ops.ConvLayer.profile_end()
    rows = ops.ConvLayer.last_profile
    agg = {}
    for r in rows:
        taps = r["k"] * r["k"]
        opos = r["B"] * r["OH"] * r["OW"]
        if r["kind"] == "ru":  # fused 1x1 -> 3x3 -> 1x1 bottleneck: C -> C/2 -> C/2 -> C
            c, m = r["cin"], r["cin"] // 2
            macs = opos * (c * m + 9 * m * m + m * c)
        elif r["kind"] == "c3":  # 3-input-channel first layers
            macs = opos * 3 * r["cout"] * taps
        else:  # ConvLayer rows carry their algorithmic MACs (live taps only; deconv = 25/4 taps per output)
            macs = r.get("alg_macs", opos * r["cin"] * r["cout"] * taps / (4 if r["kind"] == 1 else 1))
        byts = r["B"] * r["H"] * r["W"] * r["cin"] * 2 + opos * r["cout"] * (4 if r["f32"] else 2)
        if r["kind"] == "c3":
            byts = r["B"] * r["H"] * r["W"] * 3 * 4 * 2 + opos * r["cout"] * 2
        r["tflops"] = 2 * macs / r["ms"] / 1e9
        r["gbs"] = byts / r["ms"] / 1e6
        key = (r["kind"], r["cin"], r["cout"], r["k"], r["stride"], r["dil"], r["H"], r["W"], r["epi"], r["f32"], r["sq"])
        g = agg.setdefault(key, dict(n=0, ms=0.0, macs=0.0, bytes=0.0))
        g["n"] += 1
        g["ms"] += r["ms"]
        g["macs"] += macs
        g["bytes"] += byts
    print(f"total conv ms {tot:.3f} over {n} launches")
    print(f"{'kind cin->cout k s d  HxW epi f32 sq':48s} {'n':>3s} {'ms':>8s} {'share':>6s} {'TF/s':>7s} {'GB/s':>7s}")
    for key, g in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        kind, cin, cout, k, s, d, Hh, Ww, epi, f32, sq = key
        name = f"{kind if isinstance(kind, str) else ('dc' if kind else 'cv')} {cin}->{cout} k{k} s{s} d{d} {Hh}x{Ww} e{epi} {int(f32)} {int(sq)}"
        print(f"{name:48s} {g['n']:3d} {g['ms']:8.3f} {g['ms'] / tot:6.1%} {2 * g['macs'] / g['ms'] / 1e9:7.1f} {g['bytes'] / g['ms'] / 1e6:7.0f}")
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as fh:
        json.dump(dict(total_ms=tot, launches=n, rows=rows), fh)


if __name__ == "__main__":
    main()
