#!/usr/bin/env python3
"""Headline metrics of every kernel in an ncu report. usage: tools/ncu_multi.py report.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.sum", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.avg", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, rows[1]))
    for w in want:
        if w in d: print(f"{w:72s} {d[w]:>22s} {u[w]}")
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and float(d[h]) > 0.1:
            print("   stall", h.split("stalled_")[1].split("_per_issue")[0], d[h])
    print()
