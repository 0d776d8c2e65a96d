"""Text summaries of ncu output for profiles/ (the .ncu-rep files themselves stay in gpurun_out/, which is scratch).
  python tools/ncu_summaries.py rep   <file.ncu-rep>  > profiles/<name>.summary.txt      key metrics of one captured kernel
  python tools/ncu_summaries.py list  <launches.csv> '<title>' > profiles/<name>.summary.txt   launch list -> per-kernel totals
  python tools/ncu_summaries.py stalls <file.ncu-rep> [min_samples] > profiles/<name>_stalls.txt   hottest SASS lines by stall samples
"""
import csv, subprocess, sys, re, collections

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max']


def page(rep, which):
    return list(csv.reader(subprocess.run(['ncu', '-i', rep, '--page', which, '--csv'], capture_output=True, text=True).stdout.splitlines()))


def rep(path):
    rows = page(path, 'raw')
    h, u, d = rows[0], rows[1], rows[2]
    print('%-88s %s' % ('Kernel Name', d[h.index('Kernel Name')][:90]))
    for k in KEYS:
        if k in h:
            print('%-75s %-12s %s' % (k, u[h.index(k)], d[h.index(k)]))


def launches(path, title):
    rows = [r for r in csv.reader(open(path)) if r]
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    h = rows[hi]
    kn, mv, mu = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
    tot = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        name = re.sub(r'^void\s+|hd::|\(anonymous namespace\)::|<unnamed>::|\(.*$', '', r[kn])
        name = re.sub(r'\((int|bool)\)', '', name)
        t = float(r[mv].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[mu], 1.0)
        a = tot.setdefault(name, [0.0, 0])
        a[0] += t; a[1] += 1
    total = sum(v[0] for v in tot.values())
    print('# ' + title)
    print('total %.1f us' % total)
    for name, (t, c) in sorted(tot.items(), key=lambda x: -x[1][0]):
        print('%10.1f us %5.1f%% x%-3d %s' % (t, 100 * t / total, c, name))


def stalls(path, min_samples=100):
    rows = page(path, 'source')
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
    h, data = rows[hi], rows[hi + 1:]
    ix = {n: i for i, n in enumerate(h)}
    names = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
    S = lambda r: int(r[ix['# Samples']] or 0)
    print('# kernel: %s' % rows[0][1][:120])
    print('# %d SASS lines, %d warp-stall samples; lines with >= %d samples (line, samples, executed, instruction, dominant stalls)' % (
        len(data), sum(S(r) for r in data), min_samples))
    agg = {n: sum(int(r[ix[n]] or 0) for r in data) for n in names}
    print('# all lines: ' + ', '.join('%s %d' % (n[6:], v) for n, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
    for i, r in enumerate(data):
        if S(r) >= min_samples:
            dom = {n[6:]: int(r[ix[n]]) for n in names if int(r[ix[n]] or 0) * 4 >= S(r)}
            print('%5d %6d %9s  %-80s %s' % (i, S(r), r[ix['Instructions Executed']], r[ix['Source']].strip()[:80], dom))


if __name__ == '__main__':
    cmd = sys.argv[1]
    if cmd == 'rep':
        rep(sys.argv[2])
    elif cmd == 'list':
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else '')
    else:
        stalls(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 100)
