#!/usr/bin/env python
"""Per-kernel totals from an ncu `--metrics gpu__time_duration.sum --csv` launch list.
usage: scripts/launch_summary.py gpurun_out/launches_x.csv "note" > profiles/x_launch_summary.txt"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
tot = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    v *= {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'nsecond': 1e-6, 'ms': 1.0, 'msecond': 1.0}.get(r[ui], 1e-3)
    t = tot.setdefault(r[ki], [0, 0.0])
    t[0] += 1; t[1] += v
al = sum(t[1] for t in tot.values())
print('# %s' % (sys.argv[2] if len(sys.argv) > 2 else ''))
print('# kernel | launches | total ms | avg ms | share')
for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print('%-50s n=%3d total=%8.3f ms avg=%7.4f share=%5.1f%%' % (k[:50], n, ms, ms / n, 100 * ms / al))
