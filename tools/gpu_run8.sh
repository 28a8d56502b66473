set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2h_tests.log 2>&1; tail -5 gpurun_out/r2h_tests.log
CINTB200_TIMING=1 timeout 300 python tools/jk_once.py 80 > gpurun_out/r2h_jk_once_80.log 2>&1; tail -12 gpurun_out/r2h_jk_once_80.log
timeout 900 python bench.py --no-df --no-extra --e2e-tile-steps 0 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; tail -c 1500 gpurun_out/r2h_bench.json
