#!/usr/bin/env python
"""Summarise `ncu --page raw --csv` / `--page source --csv` exports: key metrics per kernel, and the instruction / stall
totals of the source page grouped into the kernel's phases by source line ranges.
usage: ncu_summary.py raw.csv [source.csv]"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmalite.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.avg.per_cycle_active",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_wait.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warp_latency_issue_stalled_not_selected.ratio",
        "smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio", "smsp__average_warp_latency_issue_stalled_no_instruction.ratio",
        "smsp__average_warp_latency_issue_stalled_branch_resolving.ratio", "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "local_load", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
names, units, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2:]
for v in vals:
    if not v or not v[0].isdigit():
        continue
    d = dict(zip(names, v))
    print("==", d.get("Kernel Name", "?")[:110])
    for k in KEYS:
        for nm in names:
            if nm == k or (k in nm and k.startswith("local")):
                print(f"  {nm:75s} {d[nm]:>16s} {units[names.index(nm)]}")
if len(sys.argv) > 2:
    rows = list(csv.reader(open(sys.argv[2])))
    h = next(i for i, r in enumerate(rows) if r and "Source" in r and any("Executed" in c for c in r))
    cols = rows[h]
    print("source columns:", cols)
