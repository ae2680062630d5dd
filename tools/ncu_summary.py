#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page raw --csv` output: one block per profiled launch with the roofline-relevant metrics."""
import csv
import sys

WANT = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("-" * 100)
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:75s} {r[i][:90]} {units[i]}")
    if 'dram__bytes_read.sum' in hdr:
        def val(name):
            i = hdr.index(name)
            v = float(r[i].replace(',', ''))
            u = units[i]
            return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
        def tval(name):
            i = hdr.index(name)
            v = float(r[i].replace(',', ''))
            return v * {'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1, 'usecond': 1e-6, 'msecond': 1e-3, 'nsecond': 1e-9, 'second': 1}.get(units[i], 1)
        tr = val('dram__bytes_read.sum') + val('dram__bytes_write.sum')
        t = tval('gpu__time_duration.sum')
        print(f"{'=> dram traffic (read+write)':75s} {tr / 1e6:.3f} MB ; {tr / t / 1e9:.1f} GB/s over {t * 1e6:.1f} us")
