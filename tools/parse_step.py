"""Aggregate an ncu gpu__time_duration launch list by kernel name."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr, data = rows[0], rows[1:]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
tot = 0.0
for r in data:
    t = float(r[vi].replace(',', '')); u = r[ui]
    t = t / 1e3 if u == 'ns' else (t * 1e3 if u == 'ms' else t)
    name = r[ki].split('(')[0].replace('void ', '').replace('hd::<unnamed>::', '').replace('<unnamed>::', '')[:70]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t; tot += t
print('total %.1f us over %d launches' % (tot, len(data)))
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%9.1f us %5.1f%%  x%-4d %s' % (t, 100 * t / tot, c, n))
