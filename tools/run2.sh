cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/r2_$name.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_$name.txt; tail -5 gpurun_out/r2_$name.txt; }
run smpl       python -m pytest tests/test_gpu_smpl.py -q -m gpu
run preproc    python -m pytest tests/test_preprocess.py -q -m gpu
run configs    python -m pytest tests/test_gpu_configs.py -q -m gpu
run nets_res   python -m pytest tests/test_gpu_nets.py -q -m gpu -k "resnet or full_window or tester"
run profsmpl   python tools/prof_smpl.py
run layers     python tools/layer_table.py
HD_SUBSAMPLE_RES=0 run layers_nosub python tools/layer_table.py
timeout 900 python bench.py --steps 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/r2_bench.json; tail -3 gpurun_out/r2_bench.err
timeout 300 python bench.py --workload smpl --steps 5 --no-cpu-baseline > gpurun_out/r2_bench_smpl.json 2>&1; head -c 300 gpurun_out/r2_bench_smpl.json; echo
HD_SPLIT=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_gemm_tc -s 2 -c 1 -f -o gpurun_out/r2_tepi_k64 python tools/prof_one.py 160 56 64 256 1 1 > gpurun_out/r2_ncu1.log 2>&1; tail -2 gpurun_out/r2_ncu1.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:smpl_lbs_tc -s 1 -c 1 -f -o gpurun_out/r2_lbs_tc python tools/prof_smpl.py 16384 > gpurun_out/r2_ncu2.log 2>&1; tail -2 gpurun_out/r2_ncu2.log
ls -la gpurun_out/*.ncu-rep
