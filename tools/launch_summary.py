#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into one row per kernel
(launches, total us, average us, share). Usage: python tools/launch_summary.py launches.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, agg = None, collections.OrderedDict()
for r in rows:
    if r and r[0] == 'ID':
        hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    d = dict(zip(hdr, r))
    if d.get('Metric Name') != 'gpu__time_duration.sum': continue
    v = float(d['Metric Value'].replace(',', ''))
    v = v / 1000.0 if d['Metric Unit'] == 'ns' else (v * 1000.0 if d['Metric Unit'] == 'ms' else v)
    a = agg.setdefault(d['Kernel Name'][:110], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print('%-112s %6s %11s %9s %6s' % ('kernel', 'n', 'total us', 'avg us', 'share'))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-112s %6d %11.1f %9.2f %5.1f%%' % (k, n, t, t / n, 100 * t / tot))
