cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_multi_gpu.py -q -m gpu -x -s > gpurun_out/r5_mgpu_test.txt 2>&1; echo "rc=$?" >> gpurun_out/r5_mgpu_test.txt; tail -25 gpurun_out/r5_mgpu_test.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r5_bench_2gpu.json 2> gpurun_out/r5_bench_2gpu.err; echo "bench2 rc=$?"
tail -c 1500 gpurun_out/r5_bench_2gpu.json; tail -3 gpurun_out/r5_bench_2gpu.err
