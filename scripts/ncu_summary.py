#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/: launch list CSV -> per-kernel mean duration and share;
.ncu-rep -> selected raw metrics per kernel."""
import csv
import subprocess
import sys
from collections import defaultdict

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__inst_executed_pipe_fp64.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_shfl.sum",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]


def launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    h = rows[0]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    d = defaultdict(list)
    for r in rows[1:]:
        d[r[ki].split("(")[0][:70]].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for k, v in d.items() if "dyb::" in k and "fp64_peak" not in k)
    print(f"{'kernel':72s} {'n':>4s} {'mean us':>10s} {'share':>7s}")
    for k, v in d.items():
        sh = sum(v) / tot if ("dyb::" in k and "fp64_peak" not in k) else float("nan")
        print(f"{k:72s} {len(v):4d} {sum(v) / len(v) / 1e3:10.1f} {sh:7.3f}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[h.index("Kernel Name")][:90])
        for w in WANT:
            if w in h:
                i = h.index(w)
                print(f"   {w:85s} {r[i]:>16s} {units[i]}")


def summary_json(rep, out_path):
    """profiles/*_ncu_summary.json: the per-kernel figures bench.py reads for `roofline.traffic`."""
    import json
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u = rows[0], rows[1]
    want = {"gpu__time_duration.sum": "duration_us", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
            "launch__registers_per_thread": "registers", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_active_pct",
            "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct", "smsp__inst_executed.sum": "warp_instructions",
            "lts__t_sector_hit_rate.pct": "l2_hit_pct"}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out = {"capture": f"ncu --set full --clock-control none --import-source on ({rep})", "kernels": {}}
    for r in rows[2:]:
        name = r[h.index("Kernel Name")].split("(")[0].replace("void ", "").replace("dyb::", "").replace("Model_", "")
        d = {}
        for k, v in want.items():
            if k in h:
                i = h.index(k)
                x = float(r[i].replace(",", ""))
                if v.startswith("dram"):
                    d[v + "_bytes"] = x * scale[u[i]]
                else:
                    d[v] = x
        d["dram_bytes"] = d.get("dram_read_bytes", 0) + d.get("dram_write_bytes", 0)
        out["kernels"][name] = d
    with open(out_path, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    if len(sys.argv) == 4 and sys.argv[1] == "--json":
        summary_json(sys.argv[2], sys.argv[3])
    else:
        for a in sys.argv[1:]:
            (launches if a.endswith(".csv") else full)(a)
