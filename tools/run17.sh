cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_ref_exec.py -q -m gpu > gpurun_out/r17_pytest_ref_exec.txt 2>&1; echo "rc=$?" >> gpurun_out/r17_pytest_ref_exec.txt; tail -30 gpurun_out/r17_pytest_ref_exec.txt
