set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-df --e2e-tile-steps 1 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err; tail -c 2500 gpurun_out/r2g_bench_n2.json; tail -5 gpurun_out/r2g_bench_n2.err
