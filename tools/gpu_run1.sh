set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_job.py tests/test_gpu_tiles.py -x -q -m gpu > gpurun_out/r2a_tests_job.log 2>&1
tail -15 gpurun_out/r2a_tests_job.log
timeout 300 python tools/quick_jk.py 80 > gpurun_out/r2a_quick_jk.log 2>&1; tail -3 gpurun_out/r2a_quick_jk.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2a_tests_parity.log 2>&1; tail -3 gpurun_out/r2a_tests_parity.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-df > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 1500 gpurun_out/r2a_bench.json
