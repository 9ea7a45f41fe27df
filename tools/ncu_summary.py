#!/usr/bin/env python3
"""Condense an `ncu --set full` report to the metrics DESIGN.md quotes (one column per captured launch).
usage: ncu -i rep.ncu-rep --page raw --csv | python tools/ncu_summary.py "title" > profiles/rNN_ncu_full_xxx.txt"""
import csv
import sys

KEEP = ["Kernel Name", "Block Size", "Grid Size", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__registers_per_thread",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
units = rows[1]
data = rows[2:]
print("# " + (sys.argv[1] if len(sys.argv) > 1 else "ncu --set full summary"))
for k in KEEP:
    if k in hdr:
        i = hdr.index(k)
        vals = [r[i][:26] for r in data if len(r) > i]
        print(f"{k:<95} {units[i]:<16} " + "  ".join(vals))
