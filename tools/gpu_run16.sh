set -x
cd $GRAFT_REPO_ROOT
CINTB200_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"jk_|tile_rowsum" --csv --log-file gpurun_out/r2p_jk_launches.csv python tools/jk_once.py 80 > gpurun_out/r2p_jk_ncu.log 2>&1; tail -2 gpurun_out/r2p_jk_ncu.log
timeout 300 python tools/time_variant.py 8 c2h6_ccpvqz > gpurun_out/r2p_time.log 2>&1; cat gpurun_out/r2p_time.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2p_tests.log 2>&1; tail -4 gpurun_out/r2p_tests.log
