#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list: mean / min duration per (kernel, grid, block).

    python tools/launch_list.py gpurun_out/launches.csv
"""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ki, vi, gi, bi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Block Size")
agg = collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault((r[ki].split("(")[0][:48], r[gi], r[bi]), []).append(float(r[vi].replace(",", "")))
print("%-48s %16s %14s %5s %10s %10s" % ("kernel", "grid", "block", "n", "mean us", "min us"))
for k, v in agg.items():
    print("%-48s %16s %14s %5d %10.1f %10.1f" % (k[0], k[1], k[2], len(v), sum(v) / len(v) / 1e3, min(v) / 1e3))
