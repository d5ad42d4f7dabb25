#!/usr/bin/env python
"""Top stall sites of one kernel from `ncu -i X.ncu-rep --page source --csv --kernel-id ...` output (SASS view)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
isrc, isamp, iexec = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")


def toi(x):
    try:
        return int(x)
    except ValueError:
        return 0


data = [(toi(r[isamp]), toi(r[iexec]), r[isrc].strip()) for r in rows[hi + 1:] if len(r) > iexec]
data = data[:len(data) // 2] if len(data) > 1 and data[0][2] == data[len(data) // 2][2] else data
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
top = sorted(enumerate(data), key=lambda x: -x[1][0])[:n]
for i, (s, e, src) in sorted(top):
    print(f"{i:5d} {s:7d} {100.0 * s / tot:5.1f}% exec {e:9d}  {src[:110]}")
