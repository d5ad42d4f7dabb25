#!/usr/bin/env python
"""Print the metrics that matter from an `ncu --page raw --csv` export (one kernel)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h, u = rows[hdr], rows[hdr + 1]
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit", "launch__grid_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
        "sm__inst_executed_pipe_fma", "sm__pipe_fma_cycles_active.avg.pct", "sm__inst_executed_pipe_lsu",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__pcsamp_warps_issue_stalled", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct",
        "l1tex__throughput.avg.pct", "lts__throughput.avg.pct", "dram__throughput.avg.pct", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.avg", "smsp__inst_executed_op_shared", "sm__pipe_alu_cycles_active", "sm__inst_executed_pipe_xu",
        "sm__pipe_fmaheavy", "sm__pipe_fmalite", "smsp__thread_inst_executed_per_inst_executed"]
for v in rows[hdr + 2:]:
    if len(v) != len(h):
        continue
    print("==", v[h.index("Kernel Name")][:100])
    for i, name in enumerate(h):
        if any(k in name for k in keys) and "not_issued" not in name:
            print(f"  {name} [{u[i]}] = {v[i]}")
