set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ip1 or deriv or large_lists" > gpurun_out/r2j_tests.log 2>&1; tail -15 gpurun_out/r2j_tests.log
timeout 1500 python bench.py --no-df --e2e-tile-steps 0 --no-check > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; tail -c 600 gpurun_out/r2j_bench.json; tail -5 gpurun_out/r2j_bench.err
