set -x
cd $GRAFT_REPO_ROOT
for b in 1 0; do CINTB200_JK_BULK=$b timeout 600 python tools/quick_jk.py 80; done > gpurun_out/r2y_jk.log 2>&1; cat gpurun_out/r2y_jk.log
