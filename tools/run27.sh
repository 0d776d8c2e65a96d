cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r27_pytest_full.txt 2>&1; echo "rc=$?" >> gpurun_out/r27_pytest_full.txt; tail -4 gpurun_out/r27_pytest_full.txt
timeout 200 python __graft_entry__.py smoke > gpurun_out/r27_smoke.txt 2>&1; tail -1 gpurun_out/r27_smoke.txt
timeout 400 python bench.py > gpurun_out/r27_bench.json 2> gpurun_out/r27_bench.err; echo "rc=$?"
python - <<PY
import json
t=[l for l in open('gpurun_out/r27_bench.json').read().splitlines() if l.startswith('{')]
d=json.loads(t[-1]); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['streaming']['value'], d['e2e']['uint8_frames']['value'], d['roofline']['frac'], d['clocks'], d.get('parity',{}).get('parity_max_rel'), d['gpu_launches'])
print({k:(v['value'],v.get('roofline',{}).get('frac')) for k,v in d['extra'].items()})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 191 -c 191 --csv --log-file gpurun_out/r27_launches_step.csv python tools/prof_step.py 2 > gpurun_out/r27_launches.log 2>&1; tail -1 gpurun_out/r27_launches.log
