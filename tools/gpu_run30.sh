set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_tests30.log 2>&1; tail -3 gpurun_out/r2_tests30.log
CINTB200_TIMING=1 timeout 600 python tools/quick_ip1.py > gpurun_out/r2_ip1.log 2> gpurun_out/r2_ip1.err; tail -1 gpurun_out/r2_ip1.log
timeout 600 python tools/quick_ip1.py 2>/dev/null | tail -1
python - <<'PY'
import re, collections
agg = collections.defaultdict(float); n = collections.Counter()
for line in open('gpurun_out/r2_ip1.err'):
    m = re.match(r"\[cintb200 timing\] (.*?)\s+([0-9.]+) ms", line)
    if m: agg[m.group(1)] += float(m.group(2)); n[m.group(1)] += 1
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]: print("%-50s %10.1f ms total over %d calls (both passes)" % (k, v, n[k]))
PY
timeout 300 python tools/time_variant.py 80 c60_ccpvdz
