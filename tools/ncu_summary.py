#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total, share, average."""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in csv.DictReader(lines):
        name = re.sub(r"^void ", "", r["Kernel Name"])
        name = re.sub(r"\(.*", "", name)
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}[r["Metric Unit"]]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':70s} {'n':>5s} {'total ms':>10s} {'share':>7s} {'avg us':>10s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:70]:70s} {v[0]:5d} {v[1] / 1e6:10.3f} {100 * v[1] / tot:6.1f}% {v[1] / v[0] / 1e3:10.1f}")
    print(f"{'total':70s} {sum(v[0] for v in agg.values()):5d} {tot / 1e6:10.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
