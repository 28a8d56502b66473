set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_job.py -x -q -m gpu > gpurun_out/r2x_tests.log 2>&1; tail -5 gpurun_out/r2x_tests.log
for b in 1 0; do CINTB200_JK_BULK=$b timeout 600 python tools/quick_jk.py 80; done > gpurun_out/r2x_jk.log 2>&1; cat gpurun_out/r2x_jk.log
