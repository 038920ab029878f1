#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: share of total time per kernel name.

    python tools/summarize_launches.py gpurun_out/launches.csv "header line" > profiles/rNN_launches_summary.txt
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    header = sys.argv[2:] if len(sys.argv) > 2 else []
    rows = []
    with open(path, newline="") as fh:
        lines = [l for l in fh if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        if r.get("Metric Unit", "ns") in ("us", "usecond"):
            v *= 1e3
        rows.append((r["Kernel Name"], v))
    agg = defaultdict(lambda: [0, 0.0])
    for name, ns in rows:
        short = re.sub(r"\(.*$", "", name)[:110]
        agg[short][0] += 1
        agg[short][1] += ns
    total = sum(v[1] for v in agg.values())
    for h in header:
        print(h)
    print("total %.1f us over %d launches" % (total / 1e3, len(rows)))
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%6.2f%% %5d x %8.1f us  %s" % (100 * ns / total, n, ns / n / 1e3, name))


if __name__ == "__main__":
    main()
