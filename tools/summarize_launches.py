#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
usage: python tools/summarize_launches.py gpurun_out/launches.csv [> profiles/xxx.md]"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r"^void ", "", name)[:90]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% |")
    print(f"| **total** | {sum(v[0] for v in agg.values())} | {tot:.1f} | 100% |")


if __name__ == "__main__":
    main(sys.argv[1])
