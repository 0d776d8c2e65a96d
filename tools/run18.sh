cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_cplan.py -q -m gpu > gpurun_out/r18_pytest_cplan.txt 2>&1; echo "rc=$?" >> gpurun_out/r18_pytest_cplan.txt; tail -40 gpurun_out/r18_pytest_cplan.txt
