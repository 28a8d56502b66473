set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "classes or sweep" > gpurun_out/r2m_tests.log 2>&1; tail -5 gpurun_out/r2m_tests.log
CHUNK_GB=8 timeout 300 python tools/profile_c60.py c2h6_ccpvqz > gpurun_out/r2m_profile_qz.txt 2>&1; head -30 gpurun_out/r2m_profile_qz.txt
CINTB200_NO_WIDE=1 CHUNK_GB=8 timeout 300 python tools/profile_c60.py c2h6_ccpvqz > gpurun_out/r2m_profile_qz_nowide.txt 2>&1; head -12 gpurun_out/r2m_profile_qz_nowide.txt
