#!/usr/bin/env python
"""ncu launch list (csv of `--metrics gpu__time_duration.sum`) of a bench run -> per-kernel table of ONE step
(from one jpeg_color_fwd_kernel launch to the next).   python tools/launch_summary.py in.csv out.md"""
import collections
import csv
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
with open(src) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ix = {h: i for i, h in enumerate(hdr)}
L = [(r[ix["Kernel Name"]], float(r[ix["Metric Value"]])) for r in rd]
starts = [i for i, (k, _) in enumerate(L) if "jpeg_color_fwd" in k]
step = L[starts[-2]:starts[-1]]
agg = collections.OrderedDict()
for k, t in step:
    name = re.sub(r"\(.*", "", k).replace("void ", "").replace("<unnamed>::", "")[:60]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(t for _, t in step)
out = ["# r01: ncu launch list of one bench step (`bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph`, workload configs[1])", "",
       "source: `profiles/r01_launches.csv` (`ncu --metrics gpu__time_duration.sum --clock-control none`), one complete step of the run",
       "(from one `jpeg_color_fwd_kernel` launch to the next).  Per-launch times under ncu are serialised and cold-cache: compare the SHARES with",
       "`bench.py`'s live CUDA-event numbers, not the absolutes.  `conv_res_kernel<EPI, ACT>`: EPI 0 linear, 1 add, 2 gate, 3 GDN, 4 IGDN,",
       "5 pixel scale (+ tensor-core up-add), 6 channel statistics; ACT 0 none, 1 ReLU, 2 PReLU.", "",
       "| kernel | launches / step | us / step | share |", "|---|---|---|---|"]
tn, tt = 0, 0.0
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if k.startswith("at::"):
        tn += n
        tt += t
        continue
    out.append(f"| `{k}` | {n} | {t / 1e3:.1f} | {100 * t / tot:.1f}% |")
out.append(f"| `(torch elementwise / copy / fill)` | {tn} | {tt / 1e3:.1f} | {100 * tt / tot:.1f}% |")
out.append(f"| total | {len(step)} | {tot / 1e3:.1f} | 100% |")
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out[-3:]))
