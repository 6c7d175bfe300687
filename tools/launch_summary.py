"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel name."""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]; ix = {k: i for i, k in enumerate(h)}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if r[ix['Metric Name']] != 'gpu__time_duration.sum':
        continue
    n = r[ix['Kernel Name']].split('(')[0]; v = float(r[ix['Metric Value']].replace(',', '')); u = r[ix['Metric Unit']]
    v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1}.get(u, 1e-6)
    agg[n][0] += 1; agg[n][1] += v
tot = sum(t for c, t in agg.values())
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print('%-50s %4d launches %9.3f ms %5.1f%%' % (n[:50], c, t, 100 * t / tot))
