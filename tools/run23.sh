cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus: $N"
timeout 500 python -m pytest tests/test_multi_gpu.py -q -m gpu -s > gpurun_out/r23_multi_gpu_test.txt 2>&1; echo "rc=$?" >> gpurun_out/r23_multi_gpu_test.txt; grep -E "bit-identical|passed|failed|rc=" gpurun_out/r23_multi_gpu_test.txt | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r23_bench_${N}gpu.json 2> gpurun_out/r23_bench_${N}gpu.err; echo "bench rc=$?"
python - <<PY
import json
t=[l for l in open('gpurun_out/r23_bench_${N}gpu.json').read().splitlines() if l.startswith('{')]
d=json.loads(t[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d.get('cuda_graph'), d['clocks'])
PY
