set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NG --no-extra --e2e-tile-steps 0 --chunk-gb $CG > gpurun_out/bench_r2_n${NG}_c$CG.json 2> gpurun_out/bench_r2_n${NG}_c$CG.err; tail -c 600 gpurun_out/bench_r2_n${NG}_c$CG.json; tail -3 gpurun_out/bench_r2_n${NG}_c$CG.err
