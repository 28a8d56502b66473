set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NG --no-extra > gpurun_out/bench_r2_n$NG.json 2> gpurun_out/bench_r2_n$NG.err; tail -c 1200 gpurun_out/bench_r2_n$NG.json; tail -3 gpurun_out/bench_r2_n$NG.err
