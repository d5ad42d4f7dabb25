#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (markdown table).
usage: python tools/summarize_launches.py launches.csv [n_steps]"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 14 and r[0].isdigit()]
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    name = re.sub(r"\(.*", "", r[4])
    name = re.sub(r"^void ", "", name)
    if len(name) > 90:
        name = name[:87] + "..."
    agg[name][0] += 1
    agg[name][1] += float(r[14].replace(",", "")) / 1e6
tot = sum(v[1] for v in agg.values())
print(f"| kernel | launches/step | ms/step | share |\n|---|---|---|---|")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n / steps:g} | {ms / steps:.3f} | {100 * ms / tot:.1f}% |")
print(f"\ntotal {tot / steps:.2f} ms/step over {len(rows) / steps:g} launches/step (ncu per-launch times are serialised, cold-cache)")
