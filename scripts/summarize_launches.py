#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total ms, share.
Usage: python scripts/summarize_launches.py <launches.csv> [--md]"""
import collections
import csv
import re
import sys


def summarize(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = row["Metric Unit"]
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(unit, 1e-6)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("me::<unnamed>::", "")
        agg[name][0] += 1
        agg[name][1] += v
    return sorted(agg.items(), key=lambda kv: -kv[1][1])


if __name__ == "__main__":
    rows = summarize(sys.argv[1])
    total = sum(v[1] for _, v in rows)
    if "--md" in sys.argv:
        print("| kernel | launches | total ms | share |\n|---|---|---|---|")
        for k, v in rows:
            if v[1] / total >= 0.0005:
                print(f"| {k[:80]} | {v[0]} | {v[1]:.2f} | {v[1] / total:.3f} |")
        print(f"| all kernels | {sum(v[0] for _, v in rows)} | {total:.2f} | 1 |")
    else:
        for k, v in rows:
            print(f"{k[:80]:80s} {v[0]:7d} {v[1]:11.3f} ms {v[1] / total:6.3f}")
        print(f"total {total:.3f} ms")
