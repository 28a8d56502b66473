set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_job.py -x -q -m gpu > gpurun_out/r2b_tests_job.log 2>&1; tail -12 gpurun_out/r2b_tests_job.log
CINTB200_TIMING=1 timeout 300 python tools/jk_once.py 80 > gpurun_out/r2b_jk_once_80.log 2>&1; cat gpurun_out/r2b_jk_once_80.log
CINTB200_TIMING=1 timeout 300 python tools/jk_once.py 16 > gpurun_out/r2b_jk_once_16.log 2>&1; tail -1 gpurun_out/r2b_jk_once_16.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"jk_|tile_rowsum" --csv --log-file gpurun_out/r2b_jk_launches.csv python tools/jk_once.py 80 > gpurun_out/r2b_jk_ncu.log 2>&1; tail -2 gpurun_out/r2b_jk_ncu.log
for pb in 16 occ 4 8 32; do CINTB200_PBLOCKS=$pb timeout 200 python tools/time_variant.py 80; done > gpurun_out/r2b_pblocks.log 2>&1; cat gpurun_out/r2b_pblocks.log
for v in _v2 _fr; do
  lib=$PWD/libcint_b200/libcint_b200$v.so
  [ -f $lib ] || continue
  CINTB200_LIB=$lib timeout 200 python tools/time_variant.py 80
  CINTB200_LIB=$lib CHUNK_GB=80 timeout 200 python tools/profile_c60.py > gpurun_out/r2b_profile$v.txt 2>&1
done > gpurun_out/r2b_variants.log 2>&1; cat gpurun_out/r2b_variants.log
CHUNK_GB=80 timeout 200 python tools/profile_c60.py > gpurun_out/r2b_profile_base.txt 2>&1; head -3 gpurun_out/r2b_profile_base.txt
