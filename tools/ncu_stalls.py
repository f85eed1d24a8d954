#!/usr/bin/env python
"""Top stall sites of one kernel of an .ncu-rep (SASS view, needs --import-source on).
    python tools/ncu_stalls.py rep.ncu-rep --launch 0 [--top 25]"""
import argparse, csv, io, subprocess
ap = argparse.ArgumentParser()
ap.add_argument("rep"); ap.add_argument("--launch", type=int, default=0); ap.add_argument("--top", type=int, default=25)
a = ap.parse_args()
out = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--launch-skip", str(a.launch), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print(rows[0][1][:80], "total samples", tot)
agg = {c: sum(int(r[ix[c]] or 0) for r in data) for c in stall_cols}
print("by reason:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
top = sorted(enumerate(data), key=lambda ir: -int(ir[1][ix["# Samples"]] or 0))[: a.top]
for i, r in top:
    n = int(r[ix["# Samples"]] or 0)
    why = sorted(((int(r[ix[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {n:7d} {100.0*n/tot:5.1f}%  {r[ix['Source']].strip()[:70]:70s} {why}")
if True:
    print("--- samples per 40-instruction block")
    blk = 40
    for b in range(0, len(data), blk):
        s = sum(int(r[ix["# Samples"]] or 0) for r in data[b:b+blk])
        ex = sum(int(r[ix["Instructions Executed"]] or 0) for r in data[b:b+blk])
        ops = {}
        for r in data[b:b+blk]:
            op = r[ix["Source"]].strip().split()[0 if not r[ix["Source"]].strip().startswith("@") else 1].split(".")[0]
            ops[op] = ops.get(op, 0) + 1
        topops = " ".join(f"{k}{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:6])
        print(f"{b:5d} {s:7d} {100.0*s/tot:5.1f}% exec {ex:10d}  {topops}")
