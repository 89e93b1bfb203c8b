#!/usr/bin/env python
"""Aggregate an ncu launch-list CSV (gpu__time_duration.sum) by kernel name and grid."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, agg = None, collections.OrderedDict()
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        agg.setdefault((d["Kernel Name"][:48], d["Grid Size"]), []).append(float(d["Metric Value"]) / 1000)
tot = 0.0
for k, v in agg.items():
    tot += sum(v)
    print(f"{k[0]:50s} {k[1]:14s} n={len(v):3d} mean={sum(v)/len(v):8.1f} us  [{' '.join(f'{x:.0f}' for x in v[:5])}]")
print(f"total {tot:.0f} us over {sum(len(v) for v in agg.values())} launches")
