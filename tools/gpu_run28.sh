set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python bench.py > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; tail -c 400 gpurun_out/bench_r2_n1.json
timeout 900 python bench.py --impl reference > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2_reference.err; tail -c 300 gpurun_out/bench_r2_reference.json
ROUND=r2 timeout 1500 bash tools/ncu_profile_r2.sh launches pipes full > gpurun_out/r2_ncu.log 2>&1; tail -5 gpurun_out/r2_ncu.log
