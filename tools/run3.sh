cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/r3_$name.txt 2>&1; echo "rc=$?" >> gpurun_out/r3_$name.txt; tail -5 gpurun_out/r3_$name.txt; }
run presplit   python -m pytest tests/test_gpu_nets.py -q -m gpu -k "presplit or conv1_from_padded"
run smpl       python -m pytest tests/test_gpu_smpl.py -q -m gpu
run configs    python -m pytest tests/test_gpu_configs.py -q -m gpu
run layers     python tools/layer_table.py
timeout 900 python bench.py --steps 5 > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r3_bench.json; tail -3 gpurun_out/r3_bench.err
HD_SPLIT=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_gemm_tc -s 2 -c 1 -f -o gpurun_out/r3_tepi_k64 python tools/prof_one.py 160 56 64 256 1 1 > gpurun_out/r3_ncu1.log 2>&1; tail -2 gpurun_out/r3_ncu1.log
