cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/sweep_chunks.py "160,640 32,640 64,640 80,640 128,640 320,640 640,640 160,320 80,320 160,160" > gpurun_out/r19_sweep.txt 2>&1; echo "rc=$?" >> gpurun_out/r19_sweep.txt; tail -14 gpurun_out/r19_sweep.txt
