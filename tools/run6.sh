# compute-sanitizer on tiny shapes of every tcgen05 / TMA kernel + a launch list of one full step + frame_chunk sweep
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
run() { t=$1; name=$2; shift; shift; echo "=== $name"; timeout $t "$@" > gpurun_out/r6_$name.txt 2>&1; echo "rc=$?" >> gpurun_out/r6_$name.txt; grep -E "ERROR SUMMARY|passed|failed|rc=" gpurun_out/r6_$name.txt | tail -4; }
run 200 smpl python -m pytest tests/test_gpu_smpl.py tests/test_gpu_configs.py -q -m gpu -x -k "smpl or c5"
run 120 profsmpl python tools/prof_smpl.py
run 600 memcheck_conv  $S --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_nets.py -q -m gpu -x -k "(presplit and shape0) or (presplit and shape10) or (conv1_from_padded and 2-24)"
run 600 memcheck_smpl  $S --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_smpl.py -q -m gpu -x -k "tensor_core_blend_path and 256"
run 300 memcheck_prep  $S --tool memcheck --print-limit 5 python -m pytest tests/test_preprocess.py -q -m gpu -x -k "run_video"
run 600 racecheck_conv $S --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_nets.py -q -m gpu -x -k "presplit and True-shape0"
run 600 synccheck_conv $S --tool synccheck --print-limit 5 python -m pytest tests/test_gpu_nets.py -q -m gpu -x -k "presplit and True-shape0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 199 -c 199 --csv --log-file gpurun_out/r6_launches_step.csv python tools/prof_step.py 2 > gpurun_out/r6_launches.log 2>&1; tail -2 gpurun_out/r6_launches.log
for fc in 128 320; do HD_FRAME_CHUNK=$fc timeout 300 python bench.py --steps 5 --no-cpu-baseline --no-extra > gpurun_out/r6_bench_fc$fc.json 2>/dev/null; head -c 220 gpurun_out/r6_bench_fc$fc.json; echo; done
