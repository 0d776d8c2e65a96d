"""Summarise an ncu launch list (gpu__time_duration) of tools/prof_resnet.py: per-shape time and TFLOP/s."""
import csv, re, sys
lf, of = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(lf)) if len(r) > 5]
hdr, data = rows[0], rows[1:]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
ops = [l.strip() for l in open(of) if l.startswith('op')]
ts = []
for r in data:
    t = float(r[vi].replace(',', '')); u = r[ui]
    t = t / 1e3 if u == 'ns' else (t * 1e3 if u == 'ms' else t)
    ts.append((r[ki], t))
conv = [t for n, t in ts if 'conv_gem' in n][1:]
tot = sum(t for _, t in ts)
print('total %.0f us/pass; conv1 %.0f; pool %.0f; convs %.0f; avgpool %.0f' % (tot, ts[0][1], ts[1][1], sum(conv), ts[-1][1]))
agg = {}
for o, t in zip(ops, conv):
    m = re.search(r'M=\s*(\d+) K=\s*(\d+) N=\s*(\d+) k(\d)x(\d) s(\d)', o); M, K, N, kh, kw, s = map(int, m.groups())
    a = agg.setdefault((M, K, N, kh, s), [0, 0.0]); a[0] += 1; a[1] += t
print('%8s %6s %6s %3s %2s  %3s %9s %8s %7s' % ('M', 'K', 'N', 'k', 's', 'cnt', 'us_total', 'us_each', 'TF/s'))
for (M, K, N, kh, s), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    bn = 64 if N <= 64 else 128
    print('%8d %6d %6d %3d %2d  %3d %9.1f %8.1f %7.1f  CTAs=%d' % (M, K, N, kh, s, c, t, t / c, 2 * M * K * N * c / t / 1e6, ((M + 127) // 128) * ((N + bn - 1) // bn)))
