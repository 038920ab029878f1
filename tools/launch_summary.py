#!/usr/bin/env python
"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list.  usage: launch_summary.py launches.csv"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if r]
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    unit = r[mu]
    us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
    name = re.sub(r"\(.*$", "", r[kn])[:100]
    agg[name][0] += 1
    agg[name][1] += us
tot = sum(v[1] for v in agg.values())
print("total %.1f us over %d launches" % (tot, sum(v[0] for v in agg.values())))
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%6.2f%% %5d x %8.1f us  %s" % (100 * us / tot, n, us / n, name))
