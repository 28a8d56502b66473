set -x
cd $GRAFT_REPO_ROOT
timeout 200 python tools/time_variant.py 80 > gpurun_out/r2e_time.log 2>&1; cat gpurun_out/r2e_time.log
timeout 200 python tools/time_variant.py 40 >> gpurun_out/r2e_time.log 2>&1; tail -1 gpurun_out/r2e_time.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2e_tests.log 2>&1; tail -5 gpurun_out/r2e_tests.log
CHUNK_GB=80 timeout 200 python tools/profile_c60.py > gpurun_out/r2e_profile.txt 2>&1; head -3 gpurun_out/r2e_profile.txt
timeout 900 python bench.py --no-df --no-extra --no-cpu --e2e-tile-steps 0 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -c 1200 gpurun_out/r2e_bench.json
