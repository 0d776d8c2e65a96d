cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for i in 1 2; do
  for e in 1 0; do
    echo "HD_SUBSAMPLE_EPI=$e"; HD_SUBSAMPLE_EPI=$e timeout 200 python tools/sweep_chunks.py "160,640" 2>&1 | tail -1
  done
done > gpurun_out/r25_ab.txt 2>&1
cat gpurun_out/r25_ab.txt
nvidia-smi --query-gpu=clocks.sm,clocks_throttle_reasons.active,power.draw --format=csv
