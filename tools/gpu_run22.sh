set -x
cd $GRAFT_REPO_ROOT
for job in c60_ccpvdz; do
  TUNE_TAG=_alt timeout 900 python tools/tune_classes.py run $job 3 > gpurun_out/r2u_tune_$job.log 2>&1; head -40 gpurun_out/r2u_tune_$job.log
done
CINTB200_TIMING=1 timeout 600 python tools/e2e_phases.py > gpurun_out/r2u_e2e.log 2>&1; grep -v "list:" gpurun_out/r2u_e2e.log | tail -60
