set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2r_tests.log 2>&1; tail -5 gpurun_out/r2r_tests.log
for ne in 0 1; do CINTB200_NO_EPITAB=$ne timeout 300 python tools/time_variant.py 8 c2h6_ccpvqz; done > gpurun_out/r2r_time.log 2>&1; cat gpurun_out/r2r_time.log
CHUNK_GB=8 timeout 300 python tools/profile_c60.py c2h6_ccpvqz > gpurun_out/r2r_profile_qz.txt 2>&1; head -8 gpurun_out/r2r_profile_qz.txt
timeout 1500 python bench.py --no-df --e2e-tile-steps 0 --no-check > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; tail -c 300 gpurun_out/r2r_bench.json; tail -3 gpurun_out/r2r_bench.err
