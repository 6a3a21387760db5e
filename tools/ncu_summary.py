#!/usr/bin/env python
"""Summarise an .ncu-rep (one `ncu --set full --import-source on` capture) as markdown:
key counters, the executed-instruction mix per warp and the top stall sites.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep "title" [warps_per_launch] > profiles/xyz.md
"""
import csv
import io
import subprocess
import sys
from collections import Counter

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fp64.sum",
        "sm__cycles_elapsed.max", "sm__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum"]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, "--csv", *args], capture_output=True, text=True).stdout


def main():
    rep, title = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw"))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(f"# {title}\n")
    print(f"Source: `{rep.split('/')[-1]}` ({len(data)} launch(es) captured with `ncu --set full --clock-control none --import-source on`).\n")
    name = data[0][hdr.index("Kernel Name")]
    print(f"Kernel: `{name}`\n")
    print("| metric | unit | " + " | ".join(f"launch {i + 1}" for i in range(len(data))) + " |")
    print("|---|---|" + "---|" * len(data))
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"| {k} | {units[i]} | " + " | ".join(r[i] for r in data) + " |")
    src = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--print-source", "sass"))))
    h = src[1]
    iS, iSm, iE = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    body = []
    for r in src[2:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) > 10 and r[0] != "Address":
            body.append(r)
    warps = float(sys.argv[3]) if len(sys.argv) > 3 else None
    mix, smp = Counter(), Counter()
    for r in body:
        t = r[iS].strip().split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        mix[op] += int(r[iE])
        smp[op] += int(r[iSm])
    tot = sum(mix.values())
    print(f"\nExecuted warp-instructions (first launch): {tot}" + (f" = {tot / warps:.1f} per warp of 32 envs" if warps else ""))
    div = warps or 1.0
    print("\n| opcode | executed" + (" per warp" if warps else "") + " | stall samples |\n|---|---|---|")
    for op, c in mix.most_common(24):
        print(f"| {op} | {c / div:.1f} | {smp[op]} |")
    print(f"\nTop stall sites ({sum(smp.values())} samples):\n")
    for r in sorted(body, key=lambda r: -int(r[iSm]))[:10]:
        print(f"- {r[iSm]} samples: `{r[iS].strip()[:100]}`")


if __name__ == "__main__":
    main()
