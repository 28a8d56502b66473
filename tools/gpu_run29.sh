set -x
cd $GRAFT_REPO_ROOT
CINTB200_TIMING=1 timeout 600 python tools/quick_ip1.py > gpurun_out/r2_ip1.log 2> gpurun_out/r2_ip1.err; tail -2 gpurun_out/r2_ip1.log
python - <<'PY'
import re, collections
agg = collections.defaultdict(float); n = collections.Counter()
for line in open('gpurun_out/r2_ip1.err'):
    m = re.match(r"\[cintb200 timing\] (.*?)\s+([0-9.]+) ms", line)
    if m: agg[m.group(1)] += float(m.group(2)); n[m.group(1)] += 1
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]): print("%-50s %10.1f ms total over %d calls (both passes)" % (k, v, n[k]))
PY
