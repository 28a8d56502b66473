set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2n_tests.log 2>&1; tail -5 gpurun_out/r2n_tests.log
CHUNK_GB=8 timeout 300 python tools/profile_c60.py c2h6_ccpvqz > gpurun_out/r2n_profile_qz.txt 2>&1; head -14 gpurun_out/r2n_profile_qz.txt
timeout 200 python tools/time_variant.py 80 > gpurun_out/r2n_time.log 2>&1; cat gpurun_out/r2n_time.log
timeout 1500 python bench.py --no-df --e2e-tile-steps 0 --no-check --no-e2e > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; tail -c 300 gpurun_out/r2n_bench.json; tail -3 gpurun_out/r2n_bench.err
