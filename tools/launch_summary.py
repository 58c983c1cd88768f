#!/usr/bin/env python
"""Aggregates an ncu launch list (gpu__time_duration.sum csv) by kernel. usage: tools_launch_summary.py csv"""
import csv, sys
from collections import OrderedDict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; i_name = hdr.index('Kernel Name'); i_val = hdr.index('Metric Value')
agg = OrderedDict()
for r in rows[1:]:
    n = r[i_name].split('(')[0].replace('void ', '')[:70]; v = float(r[i_val].replace(',', ''))
    agg.setdefault(n, [0, 0.0]); agg[n][0] += 1; agg[n][1] += v
tot = sum(v for c, v in agg.values())
print(f"{len(rows)-1} launches, {tot/1e6:.3f} ms total (cold-cache, serialised: compare shares, not absolutes)")
for n, (c, v) in agg.items():
    print(f"  {n:72s} launches={c:4d} total={v/1e6:10.3f} ms  avg={v/c/1e3:10.1f} us  share={v/tot:.4f}")
