set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "classes or sweep or ip1_dense or c2h6_large" > gpurun_out/r2l_tests.log 2>&1; tail -15 gpurun_out/r2l_tests.log
for nw in 0 1; do CINTB200_NO_WIDE=$nw timeout 300 python tools/time_variant.py 8 c2h6_ccpvqz; done > gpurun_out/r2l_time.log 2>&1; cat gpurun_out/r2l_time.log
