set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2w_tests.log 2>&1; tail -3 gpurun_out/r2w_tests.log
CINTB200_TIMING=1 timeout 600 python tools/e2e_phases.py > gpurun_out/r2w_e2e.log 2>&1; grep "step" gpurun_out/r2w_e2e.log
CINTB200_TIMING=1 timeout 900 python bench.py --no-extra --no-df --e2e-tile-steps 0 > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2w_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['s_per_step'], d['e2e']['value'])"; grep -v "list:" gpurun_out/r2w_bench.err | grep "destroy" | tail -12
