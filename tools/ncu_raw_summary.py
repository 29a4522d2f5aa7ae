#!/usr/bin/env python3
"""`ncu -i X.ncu-rep --page raw --csv` -> one JSON line per captured launch with the metrics DESIGN.md / bench.py quote."""
import csv
import json
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fmaheavy.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    head, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(head, r))
        u = dict(zip(head, units))
        out = {"kernel": d.get("Kernel Name", "")[:60]}
        for k in KEEP:
            if k in d:
                out[k] = f"{d[k]} {u[k]}".strip()
        try:
            out["_traffic"] = sum(float(d[k].replace(",", "")) * SCALE[u[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        except (KeyError, ValueError):
            pass
        print(json.dumps(out))


if __name__ == "__main__":
    main(sys.argv[1])
