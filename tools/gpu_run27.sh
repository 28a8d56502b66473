set -x
cd $GRAFT_REPO_ROOT
TUNE_TAG=_fc timeout 900 python tools/tune_classes.py run c60_ccpvdz 2 > gpurun_out/r2z_tune.log 2>&1; head -30 gpurun_out/r2z_tune.log
