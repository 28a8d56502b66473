set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2c_tests.log 2>&1; tail -8 gpurun_out/r2c_tests.log
timeout 900 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -c 3000 gpurun_out/r2c_bench.json; tail -5 gpurun_out/r2c_bench.err
