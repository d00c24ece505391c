#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of the step)."""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, mi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[mi] != 'gpu__time_duration.sum':
        continue
    name = r[ki].split('(')[0]
    t = float(r[vi].replace(',', ''))
    t = {'ns': t / 1e3, 'us': t, 'ms': t * 1e3, 's': t * 1e6}.get(r[ui], t / 1e3)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += t
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':44s} {'n':>5s} {'total_us':>10s} {'avg_us':>9s} {'share':>6s}")
for n, (c, t) in agg.items():
    print(f"{n:44s} {c:5d} {t:10.1f} {t/c:9.1f} {t/tot*100:5.1f}%")
print(f"{'TOTAL':44s} {sum(a[0] for a in agg.values()):5d} {tot:10.1f}")
