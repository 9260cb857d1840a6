#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches / total us / share,
plus the slowest individual launches.  Usage: python tools/launch_summary.py launches.csv [first_id last_id]"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else None
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else None
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        i = int(r["ID"])
        if (lo is not None and i < lo) or (hi is not None and i > hi):
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        rows.append((i, r["Kernel Name"].split("(")[0][:60], us, r.get("Grid Size", ""), r.get("Block Size", "")))
    tot = sum(r[2] for r in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for _, k, us, _, _ in rows:
        agg[k][0] += 1
        agg[k][1] += us
    print(f"total {tot:.1f} us over {len(rows)} launches")
    print("| kernel | launches | us | share |\n|---|---:|---:|---:|")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {n} | {us:.1f} | {100 * us / tot:.1f}% |")
    print("\nslowest launches:")
    for i, k, us, gs, bs in sorted(rows, key=lambda r: -r[2])[:25]:
        print(f"  #{i} {k} {us:.1f} us grid={gs} block={bs}")


if __name__ == "__main__":
    main()
