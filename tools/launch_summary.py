#!/usr/bin/env python
"""ncu launch list (csv of `--metrics gpu__time_duration.sum`) -> per-kernel table.

    python tools/launch_summary.py in.csv out.md --title "..." [--step-marker jpeg_color_fwd]

With --step-marker the table covers ONE step (from one launch of the marker kernel to the next; the last complete
one); without it every captured launch is aggregated (pipelined workloads interleave their steps)."""
import argparse
import collections
import csv
import re

ap = argparse.ArgumentParser()
ap.add_argument("src")
ap.add_argument("dst")
ap.add_argument("--title", default="ncu launch list")
ap.add_argument("--note", default="")
ap.add_argument("--step-marker", default=None)
a = ap.parse_args()
with open(a.src) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ix = {h: i for i, h in enumerate(hdr)}
unit = None
L = []
for r in rd:
    unit = r[ix["Metric Unit"]] if "Metric Unit" in ix else "ns"
    v = float(r[ix["Metric Value"]].replace(",", ""))
    L.append((r[ix["Kernel Name"]], v * (1e3 if unit in ("us", "usecond") else 1.0)))
if a.step_marker:
    starts = [i for i, (k, _) in enumerate(L) if a.step_marker in k]
    sel = L[starts[-2]:starts[-1]]
    scope = f"one complete step (from one `{a.step_marker}` launch to the next)"
else:
    sel = L
    scope = f"all {len(L)} captured launches"
agg = collections.OrderedDict()
for k, t in sel:
    name = re.sub(r"\(.*", "", k).replace("void ", "").replace("<unnamed>::", "")[:70]
    e = agg.setdefault(name, [0, 0.0])
    e[0] += 1
    e[1] += t
tot = sum(t for _, t in sel)
out = [f"# {a.title}", "", f"source: `{a.src}` (`ncu --metrics gpu__time_duration.sum --clock-control none`), {scope}.",
       "Per-launch times under ncu are serialised and cold-cache: compare the SHARES with the live CUDA-event numbers of "
       "`bench.py`, not the absolutes.", a.note, "",
       "| kernel | launches | us | share |", "|---|---|---|---|"]
tn, tt = 0, 0.0
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if k.startswith("at::") or k.startswith("std::enable_if") or "elementwise" in k or "cub::" in k:
        tn += n
        tt += t
        continue
    out.append(f"| `{k}` | {n} | {t / 1e3:.1f} | {100 * t / tot:.1f}% |")
out.append(f"| `(torch element-wise / copy / fill / reduce)` | {tn} | {tt / 1e3:.1f} | {100 * tt / tot:.1f}% |")
out.append(f"| total | {len(sel)} | {tot / 1e3:.1f} | 100% |")
open(a.dst, "w").write("\n".join(out) + "\n")
print("\n".join(out[-12:]))
