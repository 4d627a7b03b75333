#!/usr/bin/env python
"""Summarise `ncu --set full` reports into the small CSV / JSON files kept under profiles/.

    python tools/ncu_summarise.py OUT.csv REPORT.ncu-rep [REPORT2.ncu-rep ...] [--traffic-json profiles/slot_update_traffic.json]

Reads each report with `ncu -i REPORT --page raw --csv`, keeps the metrics the design document argues from
(duration, DRAM bytes, L2/L1 hit rates, pipe utilisation, occupancy, registers, stall reasons) and writes one
column per captured launch.  With --traffic-json the per-launch DRAM traffic of the k_slot_update launches is
averaged and written where bench.py reads `roofline.traffic` from.
"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__bytes.sum.per_second",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct",
    "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__maximum_warps_per_active_cycle_pct",
    "launch__registers_per_thread",
    "launch__block_size",
    "launch__grid_size",
    "launch__shared_mem_per_block_static",
    "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor",
    "smsp__inst_executed.sum",
    "smsp__cycles_active.avg",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__inst_executed_op_local_ld.sum",
    "smsp__inst_executed_op_local_st.sum",
]

UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def read_report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, launches = rows[0], rows[1], rows[2:]
    return head, units, launches


def main(argv):
    traffic_json = None
    captured = None
    if "--captured" in argv:
        i = argv.index("--captured")
        captured = argv[i + 1]
        argv = argv[:i] + argv[i + 2:]
    if "--traffic-json" in argv:
        i = argv.index("--traffic-json")
        traffic_json = argv[i + 1]
        argv = argv[:i] + argv[i + 2:]
    out_csv, reports = argv[0], argv[1:]
    cols = []  # (report, name, {metric: (unit, value)})
    for rep in reports:
        head, units, launches = read_report(rep)
        ki = head.index("Kernel Name")
        for row in launches:
            m = {h: (units[j], row[j]) for j, h in enumerate(head)}
            cols.append((rep, row[ki], m))
    if not cols:
        raise SystemExit("no launches captured in " + ", ".join(reports))
    metrics = [k for k in KEEP if any(k in c[2] for c in cols)]
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + ["launch%d" % (i + 1) for i in range(len(cols))])
        w.writerow(["Kernel Name", ""] + [c[1] for c in cols])
        w.writerow(["report", ""] + [c[0].split("/")[-1] for c in cols])
        for k in metrics:
            units = {c[2][k][0] for c in cols if k in c[2]}
            if len(units) == 1:
                w.writerow([k, units.pop()] + [c[2].get(k, ("", ""))[1] for c in cols])
            else:  # ncu scales units per kernel (Mbyte here, Gbyte there): keep each cell's own
                w.writerow([k, "(per cell)"] + [" ".join(reversed(c[2][k])) if k in c[2] else "" for c in cols])
    print("wrote", out_csv, "(%d launches, %d metrics)" % (len(cols), len(metrics)))

    if traffic_json:
        # per slot-update kernel (plain and record-sharing): average DRAM traffic per launch, keyed by the kernel's base
        # name; entries of kernels absent from these reports are kept
        try:
            with open(traffic_json) as f:
                doc = json.load(f)
            if "kernels" not in doc:
                doc = {"kernels": {}}
        except Exception:
            doc = {"kernels": {}}

        def bytes_of(c, key):
            unit, val = c[2][key]
            return float(val.replace(",", "")) * UNIT_SCALE[unit]

        found = 0
        for base in ("k_slot_update_heads_tma", "k_slot_update_heads_direct", "k_share_keys", "k_slot_update_shared", "k_slot_update",
                     "k_frame_heads", "k_resample_runs"):
            sel = [c for c in cols if c[1].replace("void ", "").strip().split("<")[0].split("(")[0] == base]
            if not sel:
                continue
            found += 1
            rd = sum(bytes_of(c, "dram__bytes_read.sum") for c in sel) / len(sel)
            wr = sum(bytes_of(c, "dram__bytes_write.sum") for c in sel) / len(sel)
            us = sum(float(c[2]["gpu__time_duration.sum"][1].replace(",", "")) for c in sel) / len(sel)
            grid = int(float(sel[0][2]["launch__grid_size"][1].replace(",", "")))
            block = int(float(sel[0][2]["launch__block_size"][1].replace(",", "")))
            doc["kernels"][base] = {
                "kernel": sel[0][1].split("(")[0].replace("void ", "").strip(),
                "source": "%s (ncu --set full --clock-control none, %d launches, bench.py config 2: 4096 tracks x 500 "
                          "slots, grid %d x %d threads)" % (out_csv, len(sel), grid, block),
                "dram_bytes_read_per_launch": round(rd),
                "dram_bytes_write_per_launch": round(wr),
                "dram_bytes_per_launch": round(rd + wr),
                "ncu_duration_us": round(us, 2),
            }
            if base in ("k_slot_update_shared", "k_slot_update"):
                doc["kernels"][base]["algorithmic_bytes_per_launch"] = 4096 * 500 * 1500
        if not found:
            raise SystemExit("no slot-update launch in the reports; %s left untouched" % traffic_json)
        doc["round"] = 2
        if captured:
            doc["captured"] = captured
        with open(traffic_json, "w") as f:
            json.dump(doc, f, indent=1)
            f.write("\n")
        print("wrote", traffic_json, list(doc["kernels"]))


if __name__ == "__main__":
    main(sys.argv[1:])
