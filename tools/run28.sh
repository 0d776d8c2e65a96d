# compute-sanitizer on the final kernels: tiny ResNet (3 frames, 72x72: odd maps 9 -> 5 -> 3) through every trunk kernel variant incl. the
# 64-wide deep-stage tiles and the epilogue-subsample conv3
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
run() { t=$1; name=$2; shift; shift; echo "=== $name"; timeout $t "$@" > gpurun_out/r28_$name.txt 2>&1; echo "rc=$?" >> gpurun_out/r28_$name.txt; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=" gpurun_out/r28_$name.txt | tail -4; }
run 170 memcheck_final  $S --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_nets.py -q -m gpu -x -k "epilogue_subsample and 72"
run 200 racecheck_final $S --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_nets.py -q -m gpu -x -k "epilogue_subsample and 72"
