# GPU validation pass: every group separately and bounded, so one failure or hang cannot hide the rest.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/r1_$name.txt 2>&1; echo "rc=$?" >> gpurun_out/r1_$name.txt; tail -6 gpurun_out/r1_$name.txt; }
run presplit   python -m pytest tests/test_gpu_nets.py -q -m gpu -k "presplit"
run conv1      python -m pytest tests/test_gpu_nets.py -q -m gpu -k "conv1_from_padded"
run nets       python -m pytest tests/test_gpu_nets.py -q -m gpu -k "not presplit and not conv1_from_padded"
run smpl       python -m pytest tests/test_gpu_smpl.py -q -m gpu
run preproc    python -m pytest tests/test_preprocess.py tests/test_golden.py -q -m gpu
run configs    python -m pytest tests/test_gpu_configs.py -q -m gpu
run smoke      python -c "import __graft_entry__ as g; g.smoke()"
run profsmpl   python tools/prof_smpl.py
HD_LBS_TC=0 run profsmpl_simt python tools/prof_smpl.py
run layers_new python tools/layer_table.py
HD_TMA_EPILOGUE=0 HD_CONV1_PLANES=0 run layers_old python tools/layer_table.py
timeout 900 python bench.py --steps 5 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; echo "bench rc=$?"
tail -c 3500 gpurun_out/r1_bench.json; tail -5 gpurun_out/r1_bench.err
HD_TMA_EPILOGUE=0 HD_CONV1_PLANES=0 timeout 600 python bench.py --steps 5 --no-cpu-baseline --graph 0 > gpurun_out/r1_bench_old.json 2> gpurun_out/r1_bench_old.err
head -c 300 gpurun_out/r1_bench_old.json; echo
timeout 600 python bench.py --steps 5 --no-cpu-baseline --graph 0 > gpurun_out/r1_bench_nograph.json 2> gpurun_out/r1_bench_nograph.err
head -c 300 gpurun_out/r1_bench_nograph.json; echo
timeout 300 python bench.py --workload smpl --steps 5 --no-cpu-baseline > gpurun_out/r1_bench_smpl.json 2>&1; head -c 400 gpurun_out/r1_bench_smpl.json; echo
timeout 300 python bench.py --workload single_frame --steps 5 --no-cpu-baseline > gpurun_out/r1_bench_c2.json 2>&1; head -c 400 gpurun_out/r1_bench_c2.json; echo
