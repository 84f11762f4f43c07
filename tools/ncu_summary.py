#!/usr/bin/env python3
"""Condenses an `ncu -i report.ncu-rep --page raw --csv` dump (2 000+ columns) into the per-launch columns the profiles/
summaries keep.  Usage: python tools/ncu_summary.py raw.csv summary.csv"""
import csv
import sys

COLS = ["ID", "Kernel Name", "launch__grid_size", "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active"]
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
idx = [hdr.index(c) if c in hdr else None for c in COLS]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    for r in rows:
        w.writerow([r[i] if i is not None and i < len(r) else "" for i in idx])
print(len(rows) - 2, "launches ->", sys.argv[2])
