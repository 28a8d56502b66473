set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ssc" > gpurun_out/r2q_tests.log 2>&1; tail -3 gpurun_out/r2q_tests.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2_reference.err; tail -c 600 gpurun_out/bench_r2_reference.json
CINTB200_TIMING=1 timeout 1800 python bench.py > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; tail -c 400 gpurun_out/bench_r2_n1.json; grep -v timing gpurun_out/bench_r2_n1.err | tail -5
