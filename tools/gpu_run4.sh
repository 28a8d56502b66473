set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_tiles.py -x -q -m gpu -k "multi_chunk or host" > gpurun_out/r2d_tests.log 2>&1; tail -4 gpurun_out/r2d_tests.log
timeout 600 python -m pytest tests/test_gpu_job.py -x -q -m gpu -k "host_tiles" >> gpurun_out/r2d_tests.log 2>&1; tail -4 gpurun_out/r2d_tests.log
CINTB200_TIMING=1 timeout 1200 python bench.py --no-df > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -c 2500 gpurun_out/r2d_bench.json; grep -v "^\[cintb200 timing\]" gpurun_out/r2d_bench.err | tail -5; grep "timing" gpurun_out/r2d_bench.err | tail -40
