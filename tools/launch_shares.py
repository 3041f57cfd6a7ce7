#!/usr/bin/env python3
"""Per-kernel share of device time from an ncu launch list (gpu__time_duration.sum). usage: tools/launch_shares.py launches.csv"""
import csv, collections, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(lines))
t = collections.defaultdict(float); c = collections.Counter()
for r in rows:
    k = r["Kernel Name"][:70]; t[k] += float(r["Metric Value"]); c[k] += 1
tot = sum(t.values())
for k, v in sorted(t.items(), key=lambda x: -x[1]):
    print(f"{100*v/tot:6.2f}%  n={c[k]:4d}  avg {v/c[k]/1e3:9.1f} us  {k}")
