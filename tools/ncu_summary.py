#!/usr/bin/env python3
"""Summarise an Nsight Compute report (read here on the CPU box) into the text kept under profiles/.
usage: ncu_summary.py <report.ncu-rep | raw_page.csv> [--json out.json]
Prints, per captured launch: duration, DRAM bytes, registers, shared memory, achieved occupancy, pipe utilisation
(FP64 / FMA / ALU / LSU), issue-slot utilisation, L1/L2 traffic, local-memory traffic and the top stall reasons."""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fp64.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_bytes.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "smsp__inst_executed_op_shared_ld.sum",
    "smsp__inst_executed_op_shared_st.sum", "smsp__inst_executed_op_global_ld.sum", "smsp__inst_executed_op_global_st.sum",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "sm__sass_thread_inst_executed_op_dfma_pred_on.sum",
    "sm__sass_thread_inst_executed_op_dmul_pred_on.sum", "sm__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__warps_eligible.avg.per_cycle_active",
]


def main():
    rep = sys.argv[1]
    if rep.endswith(".csv"):      # `ncu -i report --page raw --csv` already run where the report was taken
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")], "block": r[hdr.index("Block Size")], "grid": r[hdr.index("Grid Size")]}
        for i, h in enumerate(hdr):
            if h in KEYS:
                d[h] = (r[i], units[i])
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                d.setdefault("stalls", {})[h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")] = float(r[i].replace(",", "") or 0)
            if h.startswith("smsp__average_warp") and "latency" in h:
                d[h] = (r[i], units[i])
        out.append(d)
    for d in out:
        print(f"== {d['kernel'][:90]}  block {d['block']} grid {d['grid']}")
        for k in KEYS:
            if k in d:
                print(f"   {k:72s} {d[k][0]:>18s} {d[k][1]}")
        if "stalls" in d:
            top = sorted(d["stalls"].items(), key=lambda kv: -kv[1])[:8]
            print("   stall reasons (warps per issue-active cycle): " + ", ".join(f"{k} {v:.2f}" for k, v in top))
    if "--json" in sys.argv:
        json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
