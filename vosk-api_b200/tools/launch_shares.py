"""Summarise an ncu launch list (gpu__time_duration.sum csv): python launch_shares.py launches.csv [note...]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r:
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    if d.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = d['Kernel Name'].split('(')[0]
    v = float(d['Metric Value'].replace(',', ''))
    u = d['Metric Unit']
    v = v / 1000 if u in ('ns', 'nsecond') else v * 1000 if u in ('ms', 'msecond') else v
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
for line in sys.argv[2:]:
    print("# " + line)
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-44s launches %4d total %10.1f us avg %8.1f us share %5.1f%%" % (k[:44], n, t, t / n, 100 * t / tot))
