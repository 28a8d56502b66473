set -x
cd $GRAFT_REPO_ROOT
ROUND=r2 timeout 2400 bash tools/ncu_profile_r2.sh launches pipes full > gpurun_out/r2_ncu.log 2>&1; tail -5 gpurun_out/r2_ncu.log
