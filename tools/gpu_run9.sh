set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2i_tests.log 2>&1; tail -15 gpurun_out/r2i_tests.log
timeout 300 python tools/quick_batch.py > gpurun_out/r2i_quick_batch.log 2>&1; cat gpurun_out/r2i_quick_batch.log
CINTB200_LIST_HOST=1 timeout 300 python tools/quick_batch.py > gpurun_out/r2i_quick_batch_host.log 2>&1; cat gpurun_out/r2i_quick_batch_host.log
