cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { t=$1; name=$2; shift; shift; echo "=== $name"; timeout $t "$@" > gpurun_out/r7_$name.txt 2>&1; echo "rc=$?" >> gpurun_out/r7_$name.txt; tail -4 gpurun_out/r7_$name.txt | cut -c1-300; }
run 200 smpl    python -m pytest tests/test_gpu_smpl.py -q -m gpu -x
run 300 nets    python -m pytest tests/test_gpu_nets.py -q -m gpu -x -k "fmovie or ief or full_window or single_frame or tester or hal or conv1_from"
run 300 configs python -m pytest tests/test_gpu_configs.py tests/test_golden.py -q -m gpu -x
run 120 smoke   python -c "import __graft_entry__ as g; g.smoke()"
timeout 600 python bench.py --steps 5 > gpurun_out/r7_bench.json 2> gpurun_out/r7_bench.err; echo "bench rc=$?"
tail -c 1300 gpurun_out/r7_bench.json; tail -3 gpurun_out/r7_bench.err
HD_FAST_HEADS=0 timeout 300 python bench.py --steps 5 --no-cpu-baseline --no-extra > gpurun_out/r7_bench_slowheads.json 2>/dev/null; head -c 220 gpurun_out/r7_bench_slowheads.json; echo
run 200 layers python tools/layer_table.py
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --csv --log-file gpurun_out/r7_launches_step.csv python tools/prof_step.py 2 > gpurun_out/r7_launches.log 2>&1; tail -2 gpurun_out/r7_launches.log
HD_SPLIT=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_gemm_tc -s 2 -c 1 -f -o gpurun_out/r7_tepi_k2304 python tools/prof_one.py 640 14 256 256 3 0 > gpurun_out/r7_ncu2.log 2>&1; tail -2 gpurun_out/r7_ncu2.log
