set -x
cd $GRAFT_REPO_ROOT
for args in "1 1 0 0 2" "1 1 0 0 1" "1 0 0 0 2" "2 2 2 2 2 640"; do CINTB200_TIMING=1 timeout 120 python tools/quick_sweep1.py $args 2>&1 | tail -12; done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_sweep1_ncu.csv python tools/quick_sweep1.py 1 1 0 0 2 > /dev/null 2>&1; grep -v "^==" gpurun_out/r2_sweep1_ncu.csv | awk -F'","' '{print $5, $(NF-1), $NF}' | tail -14
