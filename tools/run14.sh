cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 600 python -m pytest tests/test_gpu_nets.py tests/test_gpu_smpl.py -x -q -m gpu > gpurun_out/r14_pytest_a.txt 2>&1; echo "rc=$?" >> gpurun_out/r14_pytest_a.txt; tail -4 gpurun_out/r14_pytest_a.txt
timeout 300 python tools/layer_table.py > gpurun_out/r14_layers.txt 2>&1; tail -32 gpurun_out/r14_layers.txt
timeout 600 python bench.py --steps 10 --no-extra --no-cpu-baseline > gpurun_out/r14_bench.json 2> gpurun_out/r14_bench.err; echo "rc=$?"
python - <<PY
import json
t=[l for l in open('gpurun_out/r14_bench.json').read().splitlines() if l.startswith('{')]
d=json.loads(t[-1]); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['frac'], d['clocks'])
PY
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r14_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/r14_pytest.txt; tail -4 gpurun_out/r14_pytest.txt
