#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"].split("(")[0]
    v = float(r["Metric Value"].replace(",", ""))
    if r.get("Metric Unit") in ("nsecond", "ns"):
        v /= 1e3
    elif r.get("Metric Unit") in ("msecond", "ms"):
        v *= 1e3
    tot[name][0] += 1
    tot[name][1] += v
total = sum(v[1] for v in tot.values())
print(f"total {total / 1e3:.2f} ms over {sum(v[0] for v in tot.values())} launches")
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{t / 1e3:9.3f} ms {100 * t / total:5.1f}%  n={n:5d}  avg {t / n:8.1f} us  {k}")
