"""profiles/r02_traffic.json from ncu captures of the CURRENT kernels (gpurun_out/*.ncu-rep): DRAM bytes per launch next to the
algorithmic bytes of the same launch.  usage: python tools/make_traffic_json.py <hbm_rep> <tensor_rep>"""
import csv, json, subprocess, sys, os


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, d = rows[0], rows[2]
    g = lambda k: float(d[hdr.index(k)].replace(',', ''))
    unit = lambda k: rows[1][hdr.index(k)]
    def bytes_of(k):
        v, u = g(k), unit(k)
        return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
    t = g('gpu__time_duration.sum') * {'ns': 1e-3, 'us': 1, 'ms': 1e3}[unit('gpu__time_duration.sum')]
    return {'kernel': d[hdr.index('Kernel Name')][:80], 'duration_us_under_ncu': t,
            'dram_bytes': int(bytes_of('dram__bytes_read.sum') + bytes_of('dram__bytes_write.sum')),
            'tensor_pipe_active_pct': g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'),
            'lts_throughput_pct': g('lts__throughput.avg.pct_of_peak_sustained_elapsed')}


hbm, ten = sys.argv[1], sys.argv[2]
out = {'source': 'ncu --set full --clock-control none, one launch each of the kernels shipped in this commit (tools/make_traffic_json.py)'}
h = raw(hbm)
M = 160 * 56 * 56
h.update({'layer': '1x1 conv3 + residual, 160 frames, 56x56, Cin=64 -> Cout=256 (M=501760, K=64, N=256), fp32 + split output, TMA epilogue',
          'algorithmic_bytes': M * 64 * 4 + M * 256 * 4 * 3})
h['achieved_gbs_under_ncu'] = h['dram_bytes'] / h['duration_us_under_ncu'] / 1e3
out['memory_bound_launch'] = h
t = raw(ten)
M = 640 * 14 * 14
t.update({'layer': '3x3 conv, 640 frames, 14x14, Cin=256 -> Cout=256 (M=125440, K=2304, N=256), pre-split fp16 A, split output, TMA epilogue',
          'algorithmic_bytes': M * 256 * 4 + M * 256 * 4 + 2304 * 256 * 4})
out['tensor_bound_launch'] = t
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'r02_traffic.json'), 'w'), indent=1)
print(json.dumps(out, indent=1))
