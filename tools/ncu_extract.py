#!/usr/bin/env python3
"""Turn `ncu -i <report>.ncu-rep --page raw --csv` into the tracked counters bench.py and the judge read.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_extract.py /tmp/raw.csv profiles/r2_rollout_counters.csv profiles/r2_rollout_counters.json [plies_per_launch]

Writes (1) a CSV with one row per kernel launch and the issue / lane / branch / stall / occupancy / DRAM columns of
the raw page (the full page has 2 386 columns), (2) a small JSON for the dominant kernel that bench.py folds into
its `roofline` object: issue-slot utilisation, active threads per issued instruction, thread instructions per ply,
DRAM bytes per launch.
"""
import csv
import json
import sys

KEEP = [
    "Kernel Name", "Grid Size", "Block Size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__thread_inst_executed_per_inst_executed.pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__sass_average_branch_targets_threads_uniform.pct",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_inst0.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum",
]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "second": 1e3,
        "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6}


def main():
    raw, out_csv, out_json = sys.argv[1:4]
    plies = float(sys.argv[4]) if len(sys.argv) > 4 else None
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [c for c in KEEP if c in hdr] + [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    idx = [hdr.index(c) for c in cols]
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(cols)
        w.writerow([units[i] for i in idx])
        for r in data:
            w.writerow([r[i] for i in idx])

    def val(r, name):
        i = hdr.index(name)
        v = float(r[i].replace(",", ""))
        return v * UNIT.get(units[i], 1.0)

    main_row = max(data, key=lambda r: val(r, "gpu__time_duration.sum"))  # the dominant launch
    inst = val(main_row, "smsp__inst_executed.sum")
    lanes = val(main_row, "smsp__thread_inst_executed_per_inst_executed.ratio")
    out = {
        "kernel": main_row[hdr.index("Kernel Name")],
        "source": out_csv,
        "duration_ms_under_ncu": val(main_row, "gpu__time_duration.sum"),
        "warp_issue_frac": val(main_row, "smsp__issue_active.avg.pct_of_peak_sustained_active") / 100.0,
        "active_threads_per_inst": lanes,
        "thread_issue_frac": val(main_row, "smsp__issue_active.avg.pct_of_peak_sustained_active") / 100.0 * lanes / 32.0,
        "warp_inst_per_launch": inst,
        "branch_uniformity_pct": val(main_row, "smsp__sass_average_branch_targets_threads_uniform.pct"),
        "warps_active_pct": val(main_row, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "dram_bytes_per_launch": val(main_row, "dram__bytes_read.sum") + val(main_row, "dram__bytes_write.sum"),
        "stalls_per_issue": {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""):
                             round(val(main_row, h), 3) for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")},
    }
    if plies:
        out["plies_per_launch"] = plies
        out["warp_inst_per_ply"] = inst / plies
        out["thread_inst_per_ply"] = inst * lanes / plies
    json.dump(out, open(out_json, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
