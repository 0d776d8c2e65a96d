cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python tools/prof_plan_roles.py AB > gpurun_out/r20_roles.txt 2>&1; echo "rc=$?" >> gpurun_out/r20_roles.txt; tail -80 gpurun_out/r20_roles.txt
