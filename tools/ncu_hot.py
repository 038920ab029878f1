#!/usr/bin/env python
"""Top stall-sample SASS lines of one kernel from an .ncu-rep (source page).  usage: ncu_hot.py rep kernel_regex [n]"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 20
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                     stdout=subprocess.PIPE, text=True).stdout.splitlines()
rows = list(csv.reader(out))
k = 0
while k < len(rows):
    if rows[k] and rows[k][0] == "Kernel Name":
        name = rows[k][1][:90]
        h = rows[k + 1]
        ia, isrc, ismp = h.index("Address"), h.index("Source"), h.index("# Samples")
        body = []
        k += 2
        while k < len(rows) and rows[k] and rows[k][0] != "Kernel Name":
            try:
                body.append((int(rows[k][ismp]), rows[k][isrc].strip()))
            except (ValueError, IndexError):
                pass
            k += 1
        tot = sum(x for x, _ in body) or 1
        print("==", name, "| samples", tot)
        for i, (s, src) in sorted(enumerate(body), key=lambda t: -t[1][0])[:n]:
            print("  %5.1f%%  #%-4d %s" % (100.0 * s / tot, i, src[:100]))
        break        # first matching launch only
    k += 1
