set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_tests33.log 2>&1; tail -5 gpurun_out/r2_tests33.log
timeout 1500 python bench.py --no-df --e2e-tile-steps 0 --no-check --no-e2e > gpurun_out/r2_bench33.json 2> gpurun_out/r2_bench33.err; tail -c 200 gpurun_out/r2_bench33.json; tail -3 gpurun_out/r2_bench33.err
