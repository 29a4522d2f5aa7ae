#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total and mean time, share."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    try:
        name, v = row["Kernel Name"], float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    unit = row.get("Metric Unit", "")
    v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v * 1e6 if unit == "s" else v
    key = name.split("(")[0][:80]
    agg[key][0] += 1
    agg[key][1] += v
tot = sum(v[1] for v in agg.values())
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
print(f"total {tot / 1e3:.2f} ms over {sum(v[0] for v in agg.values())} launches" + (f" ({tot / 1e3 / div:.2f} ms per proof, {div:g} proofs)" if div != 1 else ""))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print(f"{v[1] / 1e3 / div:9.3f} ms {100 * v[1] / tot:5.1f}% n={v[0]:5d} avg={v[1] / v[0]:9.1f} us  {k}")
