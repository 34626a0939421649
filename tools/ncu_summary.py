#!/usr/bin/env python
"""Summarise an Nsight Compute report (read here, on the CPU box) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep rows_in_launch bytes_per_row > profiles/<name>.txt
"""
import csv
import subprocess
import sys
from collections import Counter

rep, rows, bpr = sys.argv[1], float(sys.argv[2]), float(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
hdr, units = r[0], r[1]
WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "smsp__inst_executed.sum", "smsp__inst_executed_op_shared_atom.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]
for k, vals in enumerate(r[2:]):
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print(f"== launch {k}: {name}")
    m = {}
    for i, h in enumerate(hdr):
        if h in WANT:
            m[h] = (vals[i], units[i])
            print(f"  {h:72s} {vals[i]:>18s} {units[i]}")
    try:
        rd = float(m["dram__bytes_read.sum"][0]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[m["dram__bytes_read.sum"][1]]
        wr = float(m["dram__bytes_write.sum"][0]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[m["dram__bytes_write.sum"][1]]
        print(f"  -> traffic (dram read+write) = {rd + wr:.4g} B ; algorithmic = {rows * bpr:.4g} B ; ratio = {(rd + wr) / (rows * bpr):.3f}")
        inst = float(m["smsp__inst_executed.sum"][0])
        print(f"  -> thread-instructions per row = {inst * 32 / rows:.1f}")
    except Exception as e:  # noqa
        print("  (derived metrics unavailable:", e, ")")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
s = list(csv.reader(src.splitlines()))
try:
    h = s[1]
    iS, iI, iW = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
    data = [(x[iS].strip(), int(x[iI]), int(x[iW])) for x in s[2:] if len(x) > iW]
    tot, totw = sum(d[1] for d in data), sum(d[2] for d in data)
    c, cw = Counter(), Counter()
    for t, i, w in data:
        op = (t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0]
        c[op] += i
        cw[op] += w
    print("== SASS opcode mix (share of executed warp instructions, share of stall samples)")
    for op, n in c.most_common(16):
        print(f"  {op:10s} {100 * n / tot:6.2f}%   stalls {100 * cw[op] / max(totw, 1):6.2f}%")
    tma = [t for t, _, _ in data if "UBLKCP" in t or "UTMA" in t]
    print("  TMA instructions in this kernel:", sorted(set(x.split()[0] if not x.startswith('@') else x.split()[1] for x in tma)))
except Exception as e:  # noqa
    print("(source page unavailable:", e, ")")
