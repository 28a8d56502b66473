set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_tests36.log 2>&1; tail -3 gpurun_out/r2_tests36.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke36.log 2>&1; tail -3 gpurun_out/r2_smoke36.log
timeout 1500 python bench.py > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; tail -c 300 gpurun_out/bench_r2_n1.json; tail -3 gpurun_out/bench_r2_n1.err
