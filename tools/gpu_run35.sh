set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_tests35.log 2>&1; tail -3 gpurun_out/r2_tests35.log
for args in "1 1 0 0 2" "2 2 2 2 2 640" "3 3 3 3 2 100"; do timeout 120 python tools/quick_sweep1.py $args 2>/dev/null | tail -1; done
timeout 300 python tools/quick_ip1.py 2>/dev/null | tail -1
timeout 300 python tools/time_variant.py 8 c2h6_ccpvqz
