cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 600 python -m pytest tests/test_gpu_nets.py tests/test_golden.py -x -q -m gpu > gpurun_out/r15_pytest_a.txt 2>&1; echo "rc=$?" >> gpurun_out/r15_pytest_a.txt; tail -4 gpurun_out/r15_pytest_a.txt
timeout 300 python tools/layer_table.py > gpurun_out/r15_layers.txt 2>&1; head -8 gpurun_out/r15_layers.txt | grep -v "^ "; grep -E "2007040|2304   256 3 1|501760    64   256 1 1   1" gpurun_out/r15_layers.txt
timeout 600 python bench.py --steps 10 --no-extra --no-cpu-baseline > gpurun_out/r15_bench.json 2> gpurun_out/r15_bench.err; echo "rc=$?"
python - <<PY
import json
t=[l for l in open('gpurun_out/r15_bench.json').read().splitlines() if l.startswith('{')]
d=json.loads(t[-1]); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['frac'], d['clocks'])
PY
HD_SPLIT=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_gemm_tc -s 2 -c 1 -f -o gpurun_out/r15_k2304 python tools/prof_one.py 640 14 256 256 3 0 > gpurun_out/r15_ncu2.log 2>&1; tail -2 gpurun_out/r15_ncu2.log
HD_SPLIT=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_gemm_tc -s 2 -c 1 -f -o gpurun_out/r15_k64 python tools/prof_one.py 160 56 64 256 1 1 > gpurun_out/r15_ncu1.log 2>&1; tail -2 gpurun_out/r15_ncu1.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --csv --log-file gpurun_out/r15_launches_step.csv python tools/prof_step.py 2 > gpurun_out/r15_launches.log 2>&1; tail -2 gpurun_out/r15_launches.log
