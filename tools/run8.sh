cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { t=$1; name=$2; shift; shift; echo "=== $name"; timeout $t "$@" > gpurun_out/r8_$name.txt 2>&1; echo "rc=$?" >> gpurun_out/r8_$name.txt; tail -4 gpurun_out/r8_$name.txt | cut -c1-300; }
run 400 nets    python -m pytest tests/test_gpu_nets.py -q -m gpu -x
HD_TEPI_SLIM=0 run 400 nets_noslim python -m pytest tests/test_gpu_nets.py -q -m gpu -x -k "presplit or resnet or full_window"
run 300 configs python -m pytest tests/test_gpu_configs.py tests/test_golden.py tests/test_gpu_smpl.py -q -m gpu -x
timeout 600 python bench.py --steps 5 > gpurun_out/r8_bench.json 2> gpurun_out/r8_bench.err; echo "bench rc=$?"
tail -c 1300 gpurun_out/r8_bench.json; tail -3 gpurun_out/r8_bench.err
HD_TEPI_SLIM=0 timeout 300 python bench.py --steps 5 --no-cpu-baseline --no-extra > gpurun_out/r8_bench_noslim.json 2>/dev/null; head -c 220 gpurun_out/r8_bench_noslim.json; echo
run 200 layers python tools/layer_table.py
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --csv --log-file gpurun_out/r8_launches_step.csv python tools/prof_step.py 2 > gpurun_out/r8_launches.log 2>&1; tail -2 gpurun_out/r8_launches.log
