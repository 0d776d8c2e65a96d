cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_nets.py tests/test_golden.py tests/test_gpu_cplan.py tests/test_ref_exec.py -x -q -m gpu > gpurun_out/r24_pytest_a.txt 2>&1; echo "rc=$?" >> gpurun_out/r24_pytest_a.txt; tail -15 gpurun_out/r24_pytest_a.txt
timeout 200 python tools/layer_table.py > gpurun_out/r24_layers.txt 2>&1; head -12 gpurun_out/r24_layers.txt
timeout 300 python bench.py --steps 8 --no-extra --no-cpu-baseline > gpurun_out/r24_bench.json 2> gpurun_out/r24_bench.err; echo "rc=$?"
python - <<PY
import json
t=[l for l in open('gpurun_out/r24_bench.json').read().splitlines() if l.startswith('{')]
d=json.loads(t[-1]); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['frac'], d['clocks'], d['gpu_launches'])
PY
