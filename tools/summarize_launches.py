#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table
(total time, share, launches per kernel name).  Usage: summarize_launches.py launches.csv [title]"""
import csv
import re
import sys
from collections import defaultdict


def main(path, title="ncu launch list"):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        rows.append((r["Kernel Name"], v * scale))
    agg = defaultdict(lambda: [0.0, 0])
    for name, ms in rows:
        name = re.sub(r"\(.*$", "", name).strip()
        agg[name][0] += ms
        agg[name][1] += 1
    total = sum(v[0] for v in agg.values())
    print(f"# {title}\n")
    print("| total ms | share | launches | avg us | kernel |\n|---|---|---|---|---|")
    for name, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if ms / total < 0.0005:
            continue
        print(f"| {ms:.2f} | {100 * ms / total:.1f}% | {n} | {1e3 * ms / n:.1f} | `{name[:110]}` |")
    print(f"\nTotal {total:.1f} ms over {len(rows)} launches.")


if __name__ == "__main__":
    main(*sys.argv[1:])
