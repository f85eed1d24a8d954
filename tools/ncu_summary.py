#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small markdown table for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--out profiles/x.md] [--title "..."]
"""
import argparse
import csv
import io
import subprocess

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("launch__occupancy_limit_shared_mem", "occ lim smem"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor hmma %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor inst"),
    ("sm__inst_executed_pipe_uniform.sum", "uniform inst"),
    ("smsp__inst_executed.sum", "inst"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "st sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "st requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "ld sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "ld requests"),
]


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    return hdr, units, data


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--out")
    ap.add_argument("--title", default="")
    ap.add_argument("--grep", default="", help="also list every metric whose name contains this substring")
    a = ap.parse_args()
    hdr, units, data = load(a.rep)
    col = {h: i for i, h in enumerate(hdr)}
    lines = []
    if a.title:
        lines += [f"# {a.title}", ""]
    lines += [f"source: `{a.rep}` (`ncu --set full --clock-control none`), one column per captured launch", ""]
    names = [r[col["Kernel Name"]] for r in data]
    lines += ["| metric | unit | " + " | ".join(f"#{i} {n[:28]}" for i, n in enumerate(names)) + " |",
              "|---|---|" + "---|" * len(names)]
    mets = list(METRICS)
    if a.grep:
        mets += [(h, h) for h in hdr if a.grep in h and h not in dict(METRICS)]
    for m, label in mets:
        if m not in col:
            continue
        i = col[m]
        lines.append(f"| {label} (`{m}`) | {units[i]} | " + " | ".join(r[i] for r in data) + " |")
    txt = "\n".join(lines) + "\n"
    if a.out:
        with open(a.out, "w") as fh:
            fh.write(txt)
    print(txt)


if __name__ == "__main__":
    main()
