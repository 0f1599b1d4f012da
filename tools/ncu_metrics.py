#!/usr/bin/env python
"""Key metrics of the last profiled launch in an .ncu-rep (issue/stall/pipe/memory) - usage: ncu_metrics.py REPORT [kernel_regex]"""
import csv, io, subprocess, sys
cmd = ["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"]
if len(sys.argv) > 2:
    cmd += ["--kernel-name", f"regex:{sys.argv[2]}"]
rows = list(csv.reader(io.StringIO(subprocess.run(cmd, capture_output=True, text=True).stdout)))
h, v = rows[0], rows[-1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
keys += [n for n in h if n.startswith("smsp__average_warps_issue_stalled") and n.endswith("per_issue_active.ratio")]
for k in keys:
    if k in h:
        print(f"{k:92s} {v[h.index(k)]}")
