cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { t=$1; name=$2; shift; shift; echo "=== $name"; timeout $t "$@" > gpurun_out/r9_$name.txt 2>&1; echo "rc=$?" >> gpurun_out/r9_$name.txt; tail -4 gpurun_out/r9_$name.txt | cut -c1-300; }
run 300 stream  python -m pytest tests/test_gpu_configs.py -q -m gpu -x -k "stream or host_paths or c3_predict_host"
run 900 full    python -m pytest tests -q -m gpu -x
timeout 600 python bench.py > gpurun_out/r9_bench.json 2> gpurun_out/r9_bench.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/r9_bench.json; tail -3 gpurun_out/r9_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r9_bench_ref.json 2> gpurun_out/r9_bench_ref.err; echo "ref rc=$?"; head -c 700 gpurun_out/r9_bench_ref.json
