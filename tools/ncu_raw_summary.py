#!/usr/bin/env python
"""Compact per-kernel summary of an `ncu -i X.ncu-rep --page raw --csv` dump (the .ncu-rep itself is too large to commit).
usage: python tools/ncu_raw_summary.py raw.csv > summary.txt"""
import csv
import sys

WANT = ["Grid Size", "Block Size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
        "sm__inst_executed_pipe_lsu.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "lts__t_sector_hit_rate.pct",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print("==", r[idx["Kernel Name"]][:150])
    for w in WANT:
        if w in idx and r[idx[w]] not in ("", "n/a"):
            print(f"   {w:82s} {r[idx[w]]:>18s} {units[idx[w]]}")
    rd, wr = idx.get("dram__bytes_read.sum"), idx.get("dram__bytes_write.sum")
