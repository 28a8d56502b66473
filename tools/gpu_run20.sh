set -x
cd $GRAFT_REPO_ROOT
timeout 300 python tools/time_variant.py 8 c2h6_ccpvqz > gpurun_out/r2s_time.log 2>&1; cat gpurun_out/r2s_time.log
for job in c60_ccpvdz df c2h6_ccpvqz c2h6_ccpvtz; do
  timeout 900 python tools/tune_classes.py run $job 2 > gpurun_out/r2s_tune_$job$TUNE_TAG.log 2>&1; head -12 gpurun_out/r2s_tune_$job$TUNE_TAG.log
done
