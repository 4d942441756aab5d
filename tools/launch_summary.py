#!/usr/bin/env python3
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and per-launch series."""
import collections, csv, sys
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = list(csv.DictReader(lines))
def us(r):
    v = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']
    return v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v * 1e6 if u == 's' else v
agg = collections.OrderedDict()
for r in rows:
    name = r['Kernel Name'].split('(')[0]
    a = agg.setdefault(name, [0, 0.0, 0.0]); a[0] += 1; a[1] += us(r); a[2] = max(a[2], us(r))
tot = sum(a[1] for a in agg.values())
print(f"{len(rows)} launches, {tot/1e3:.3f} ms total")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:22s} n={a[0]:4d} total={a[1]/1e3:9.3f} ms  max={a[2]:9.1f} us  share={100*a[1]/tot:5.1f}%")
if len(sys.argv) > 2:
    for nm in sys.argv[2:]:
        s = [(r['Grid Size'], r['Block Size'], round(us(r), 1)) for r in rows if r['Kernel Name'].startswith(nm)]
        print(nm, [(g.strip('()').split(',')[0], b.strip('()').split(',')[0], t) for g, b, t in s][:160])
