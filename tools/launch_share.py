#!/usr/bin/env python
"""Share of one frame step per kernel from an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py
(profiles/rNN_launch_list_summary.txt).  Times are cold-cache and serialised, so only the shares are comparable with
the live CUDA-event figures of bench.py (`roofline.stage_ms`).

    python tools/launch_share.py gpurun_out/launches_r01.csv "<command the list was taken with>" > profiles/r01_launch_list_summary.txt
"""
import collections
import csv
import sys

STEP_KERNELS = ["k_indicator_bounds", "k_share_keys", "k_slot_update_heads_direct", "k_slot_update_shared", "k_slot_update<",
                "k_slot_update_repair", "k_resample_block", "k_resample_small", "k_estimate"]

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].split("(")[0].replace("void ", "").strip()
    agg.setdefault(name, []).append(float(r[vi].replace(",", "")))
cmd = sys.argv[2] if len(sys.argv) > 2 else "python bench.py"
print(f"ncu --metrics gpu__time_duration.sum --clock-control none, {cmd}")
print("per-launch device time (cold-cache, serialised); median = steady state (the first ~15 frames after each of the\n"
      "bench's three resets still carry more distinct records)\n")
print("%-44s %9s %12s %14s" % ("kernel", "launches", "mean ns", "median ns"))
steady = {}
for k, v in agg.items():
    steady[k] = sorted(v)[len(v) // 2]
    print("%-44s %9d %12.1f %14.1f" % (k[:44], len(v), sum(v) / len(v), steady[k]))
step = [(k, t) for k, t in steady.items() if any(k.startswith(p.rstrip("<")) and (p[-1] != "<" or k.startswith(p)) for p in STEP_KERNELS)
        and len(agg[k]) >= 8]
tot = sum(t for _, t in step)
print("\nshare of one frame step (medians of the per-step kernels):")
for k, t in step:
    print("  %-42s %8.1f us  %5.1f %%" % (k[:42], t / 1e3, 100 * t / tot))
print("  %-42s %8.1f us" % ("total", tot / 1e3))
