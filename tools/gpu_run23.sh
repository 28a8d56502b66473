set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2v_tests.log 2>&1; tail -3 gpurun_out/r2v_tests.log
timeout 300 python tools/time_variant.py 80 c60_ccpvdz > gpurun_out/r2v_time.log 2>&1; cat gpurun_out/r2v_time.log
CINTB200_TIMING=1 timeout 900 python bench.py --no-extra --no-df --e2e-tile-steps 0 > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; tail -c 1500 gpurun_out/r2v_bench.json; grep -v "list:" gpurun_out/r2v_bench.err | grep timing | tail -50
