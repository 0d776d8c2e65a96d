cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { t=$1; name=$2; shift; shift; echo "=== $name"; timeout $t "$@" > gpurun_out/r8_$name.txt 2>&1; echo "rc=$?" >> gpurun_out/r8_$name.txt; tail -4 gpurun_out/r8_$name.txt | cut -c1-300; }
run 400 nets    python -m pytest tests/test_gpu_nets.py -q -m gpu -x
run 300 configs python -m pytest tests/test_gpu_configs.py tests/test_golden.py tests/test_gpu_smpl.py -q -m gpu -x
timeout 600 python bench.py --steps 5 > gpurun_out/r8_bench.json 2> gpurun_out/r8_bench.err; echo "bench rc=$?"
tail -c 1300 gpurun_out/r8_bench.json; tail -3 gpurun_out/r8_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --csv --log-file gpurun_out/r8_launches_step.csv python tools/prof_step.py 2 > gpurun_out/r8_launches.log 2>&1; tail -2 gpurun_out/r8_launches.log
