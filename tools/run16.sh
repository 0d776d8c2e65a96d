cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 500 python -m pytest tests -x -q -m gpu > gpurun_out/r16_pytest_full.txt 2>&1; echo "rc=$?" >> gpurun_out/r16_pytest_full.txt; tail -4 gpurun_out/r16_pytest_full.txt
timeout 200 python __graft_entry__.py smoke > gpurun_out/r16_smoke.txt 2>&1; tail -2 gpurun_out/r16_smoke.txt
timeout 400 python bench.py > gpurun_out/r16_bench.json 2> gpurun_out/r16_bench.err; echo "rc=$?"
python - <<PY
import json
t=[l for l in open('gpurun_out/r16_bench.json').read().splitlines() if l.startswith('{')]
d=json.loads(t[-1]); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['frac'], d['clocks'], d.get('parity'))
PY
timeout 200 python tools/layer_table.py > gpurun_out/r16_layers.txt 2>&1; head -4 gpurun_out/r16_layers.txt
