cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus: $N"
if [ "$N" -le 2 ]; then
  timeout 600 python -m pytest tests/test_multi_gpu.py -q -m gpu -x -s > gpurun_out/r10_mgpu_test.txt 2>&1; echo "rc=$?" >> gpurun_out/r10_mgpu_test.txt; tail -5 gpurun_out/r10_mgpu_test.txt
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r10_bench_${N}gpu.json 2> gpurun_out/r10_bench_${N}gpu.err; echo "bench$N rc=$?"
tail -c 1800 gpurun_out/r10_bench_${N}gpu.json; tail -3 gpurun_out/r10_bench_${N}gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/r10_ref_${N}gpu.json 2> gpurun_out/r10_ref_${N}gpu.err; echo "ref rc=$?"; head -c 300 gpurun_out/r10_ref_${N}gpu.json
