cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus: $N"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r13_bench_${N}gpu.json 2> gpurun_out/r13_bench_${N}gpu.err
echo "bench$N rc=$?"; tail -3 gpurun_out/r13_bench_${N}gpu.err
python - <<PY
import json
t=[l for l in open('gpurun_out/r13_bench_${N}gpu.json').read().splitlines() if l.startswith('{')]
d=json.loads(t[-1]); print($N, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d.get('parity'), d['clocks'])
PY
