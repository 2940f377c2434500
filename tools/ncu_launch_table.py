#!/usr/bin/env python3
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: python tools/ncu_launch_table.py launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    a = agg.setdefault(r[kn].split("(")[0][-60:], [0, 0.0])
    a[0] += 1
    a[1] += float(r[mv].replace(",", "")) * scale.get(r[mu], 1.0)
total = sum(t for _, t in agg.values())
print(f"{'total ms':>12} {'launches':>8} {'ms/launch':>10} {'share':>6}  kernel")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:12.3f} {c:8d} {t / c:10.4f} {100 * t / total:5.1f}%  {k}")
