#!/bin/bash
# Run on the GPU box (gpurun).  usage: tools/ncu_profile.sh [launches] [full]
#   launches : every launch of one whole C60 job with its duration -> gpurun_out/launches_rN.csv (+ per-kernel summary)
#   full     : `--set full` captures of representative kernels; exported to CSV on the box (the .ncu-rep files
#              with imported source are ~25 MB each and gpurun only brings back 64 MiB)
ROUND=${ROUND:-r1}
mkdir -p gpurun_out
cat > /tmp/onejob.py <<'PY'
import sys
sys.path.insert(0, '.')
import libcint_b200 as cb
atm, bas, env = cb.load_fixture("c60_ccpvdz")
ctx = cb.Context(atm, bas, env)
ctx.set_schwarz_threshold(0.0)                     # the bounds are computed once per context, outside bench.py's timed steps:
                                                   # keep the launch list = the launches of one timed step (nothing is screened on C60)
ctx.lib.cintb200_debug_profile(ctx.handle, 1)      # single stream: launches serialised like the timed profile pass
st = ctx.all_unique(chunk_bytes=80 << 30)
print("gpu ms", st[7], "launches", st[4])
import numpy as np, os
np.save(os.path.join("gpurun_out", "launch_rows.npy"), ctx.launch_rows())
PY
for what in "$@"; do
case $what in
launches)
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$ROUND.csv python /tmp/onejob.py > gpurun_out/launches_$ROUND.log 2>&1
  ;;
full)
  capture() {   # name, mangled-regex, skip
    ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$2 -s $3 -c 1 -f -o /tmp/prof_$1 python /tmp/onejob.py > gpurun_out/prof_$1_$ROUND.log 2>&1
    ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/prof_$1_${ROUND}_raw.csv 2>/dev/null
    ncu -i /tmp/prof_$1.ncu-rep --page source --csv > gpurun_out/prof_$1_${ROUND}_source.csv 2>/dev/null
    ncu -i /tmp/prof_$1.ncu-rep --page details > gpurun_out/prof_$1_${ROUND}_details.txt 2>/dev/null
  }
  capture reg_psps eri_reg_kernelILi1ELi0ELi1ELi0ELi2ELi2E 8
  capture coop_dpdp eri_coop_kernelILi2ELi1ELi2ELi1E 8
  capture reg_sssp eri_reg_kernelILi0ELi0ELi1ELi0ELi4ELi2E 8
  ;;
esac
done
ls -la gpurun_out
