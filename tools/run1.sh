set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 600 python -m pytest tests/test_gpu_nets.py -x -q -m gpu -k "presplit" 2>&1 | tail -15 > gpurun_out/r1_presplit.txt
cat gpurun_out/r1_presplit.txt
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r1_pytest.txt
cat gpurun_out/r1_pytest.txt
timeout 300 python tools/layer_table.py > gpurun_out/r1_layers_tma.txt 2>&1
HD_TMA_EPILOGUE=0 timeout 300 python tools/layer_table.py > gpurun_out/r1_layers_old.txt 2>&1
timeout 600 python bench.py --steps 5 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
tail -c 3000 gpurun_out/r1_bench.json; tail -5 gpurun_out/r1_bench.err
HD_TMA_EPILOGUE=0 timeout 600 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/r1_bench_old.json 2> gpurun_out/r1_bench_old.err
head -c 400 gpurun_out/r1_bench_old.json
