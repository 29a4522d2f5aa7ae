#!/usr/bin/env python3
"""Condense `ncu -i X.ncu-rep --page raw --csv` into one JSON line per captured launch with the metrics DESIGN.md cites."""
import csv
import json
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "sm__icc_requests_lookup_hit.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "local_load_bytes", "smsp__cycles_active.avg"]
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
names, units = rows[hdr], rows[hdr + 1]
for r in rows[hdr + 2:]:
    if len(r) != len(names):
        continue
    d = {"kernel": r[names.index("Kernel Name")][:90]}
    for k in KEEP:
        if k in names:
            j = names.index(k)
            d[k + (" [" + units[j] + "]" if units[j] else "")] = r[j]
    print(json.dumps(d))
