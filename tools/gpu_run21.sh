set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2t_tests.log 2>&1; tail -5 gpurun_out/r2t_tests.log
(timeout 300 python tools/time_variant.py 80 c60_ccpvdz; timeout 300 python tools/time_variant.py 8 c2h6_ccpvqz; timeout 300 python tools/time_variant.py 8 c2h6_ccpvtz) > gpurun_out/r2t_time.log 2>&1; cat gpurun_out/r2t_time.log
for job in c60_ccpvdz df c2h6_ccpvqz; do
  TUNE_TAG=_2 timeout 900 python tools/tune_classes.py run $job 2 > gpurun_out/r2t_tune_$job.log 2>&1; head -6 gpurun_out/r2t_tune_$job.log
done
timeout 1500 python bench.py --no-df --e2e-tile-steps 0 --no-check > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; tail -c 300 gpurun_out/r2t_bench.json; tail -3 gpurun_out/r2t_bench.err
