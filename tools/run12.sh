cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r12_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/r12_pytest.txt; tail -6 gpurun_out/r12_pytest.txt
timeout 600 python bench.py --steps 10 --no-extra --no-cpu-baseline > gpurun_out/r12_bench_drop.json 2> gpurun_out/r12_bench_drop.err; echo "rc=$?"
HD_DROP_DEAD_FP32=0 timeout 600 python bench.py --steps 10 --no-extra --no-cpu-baseline > gpurun_out/r12_bench_keep.json 2> gpurun_out/r12_bench_keep.err; echo "rc=$?"
python - <<PY
import json
for f in ['drop','keep']:
    t=[l for l in open('gpurun_out/r12_bench_%s.json'%f).read().splitlines() if l.startswith('{')]
    d=json.loads(t[-1]); print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['parity']['parity_max_rel'], d['clocks'])
PY
timeout 300 python tools/layer_table.py > gpurun_out/r12_layers.txt 2>&1; tail -70 gpurun_out/r12_layers.txt
