cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus: $N"
timeout 300 python -m pytest tests/test_gpu_configs.py -q -m gpu -x -k "graph or stream" > gpurun_out/r11_graph_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r11_graph_tests.txt; tail -4 gpurun_out/r11_graph_tests.txt
for n in 1 $N; do
  if [ "$n" = "1" ]; then timeout 600 python bench.py --steps 5 --no-extra > gpurun_out/r11_bench_1gpu.json 2> gpurun_out/r11_bench_1gpu.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r11_bench_${n}gpu.json 2> gpurun_out/r11_bench_${n}gpu.err; fi
  echo "bench$n rc=$?"; python - <<PY
import json
t=[l for l in open('gpurun_out/r11_bench_${n}gpu.json').read().splitlines() if l.startswith('{')]
d=json.loads(t[-1]); print($n, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d.get('parity'), d.get('cuda_graph'), d['clocks'])
PY
done
