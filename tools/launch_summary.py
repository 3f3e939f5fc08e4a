#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of the captured forward)."""
import collections
import csv
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0][:70]
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':72s} {'launches':>8s} {'total us':>10s} {'share':>7s} {'avg us':>8s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:72s} {v[0]:8d} {v[1] / 1e3:10.1f} {v[1] / tot * 100:6.1f}% {v[1] / v[0] / 1e3:8.1f}")
    print(f"{'total':72s} {sum(v[0] for v in agg.values()):8d} {tot / 1e3:10.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
