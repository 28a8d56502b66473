set -x
cd $GRAFT_REPO_ROOT
for ng in 0 1; do for nm in c60_ccpvdz c2h6_ccpvdz c2h6_ccpvtz c2h6_631g; do CINTB200_NO_GRAPH=$ng timeout 200 python tools/time_variant.py 80 $nm; done; done > gpurun_out/r2f_time.log 2>&1; cat gpurun_out/r2f_time.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2f_tests.log 2>&1; tail -5 gpurun_out/r2f_tests.log
