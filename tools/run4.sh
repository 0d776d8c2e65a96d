cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { t=$1; name=$2; shift; shift; echo "=== $name"; timeout $t "$@" > gpurun_out/r4_$name.txt 2>&1; echo "rc=$?" >> gpurun_out/r4_$name.txt; tail -4 gpurun_out/r4_$name.txt | cut -c1-300; }
run 150 presplit   python -m pytest tests/test_gpu_nets.py -q -m gpu -x -k "presplit or conv1_from_padded"
if ! grep -q "rc=0" gpurun_out/r4_presplit.txt; then echo "presplit failed: stopping"; exit 0; fi
HD_TEPI_MAXK=256 run 200 nets256 python -m pytest tests/test_gpu_nets.py -q -m gpu -x -k "resnet or full_window"
run 200 nets_allk  python -m pytest tests/test_gpu_nets.py -q -m gpu -x -k "resnet or full_window"
run 200 smpl       python -m pytest tests/test_gpu_smpl.py -q -m gpu
run 300 configs    python -m pytest tests/test_gpu_configs.py -q -m gpu -x
HD_TEPI_MAXK=256 run 200 layers256 python tools/layer_table.py
run 200 layers_allk python tools/layer_table.py
run 120 profsmpl   python tools/prof_smpl.py
timeout 600 python bench.py --steps 5 > gpurun_out/r4_bench.json 2> gpurun_out/r4_bench.err; echo "bench rc=$?"
tail -c 1200 gpurun_out/r4_bench.json; tail -3 gpurun_out/r4_bench.err
HD_TEPI_MAXK=256 timeout 300 python bench.py --steps 5 --no-cpu-baseline --no-extra > gpurun_out/r4_bench256.json 2> gpurun_out/r4_bench256.err; head -c 300 gpurun_out/r4_bench256.json; echo
HD_SPLIT=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_gemm_tc -s 2 -c 1 -f -o gpurun_out/r4_tepi_k64 python tools/prof_one.py 160 56 64 256 1 1 > gpurun_out/r4_ncu1.log 2>&1; tail -2 gpurun_out/r4_ncu1.log
