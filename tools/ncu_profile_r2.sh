#!/bin/bash
# Run on the GPU box (gpurun).  usage: ROUND=r2 tools/ncu_profile_r2.sh [launches] [pipes] [full]
#   launches : every launch of ONE C60 job incl. the J/K consumer kernels, with its duration -> gpurun_out/launches_$ROUND.csv
#   pipes    : per-launch FP64-pipe / occupancy counters of one plain C60 pass -> gpurun_out/pipes_$ROUND.csv
#   full     : `--set full` captures of representative kernels, exported to CSV / text on the box
ROUND=${ROUND:-r2}
export CINTB200_NO_GRAPH=1          # kernel launches stay individual launches under the profiler
mkdir -p gpurun_out
cat > /tmp/onejob.py <<'PY'
import sys, os
sys.path.insert(0, '.')
import numpy as np
import libcint_b200 as cb
mode = sys.argv[1] if len(sys.argv) > 1 else "plain"
atm, bas, env = cb.load_fixture("c60_ccpvdz")
ctx = cb.Context(atm, bas, env)
ctx.set_schwarz_threshold(0.0)                     # the bounds are computed once per context, outside bench.py's timed steps
ctx.lib.cintb200_debug_profile(ctx.handle, 1)      # single stream: launches serialised like the timed profile pass
if mode == "jk":
    _, _, D, _ = cb.job_weights(840)
    ctx.lib.cintb200_debug_profile(ctx.handle, 0)
    vj, vk, st = ctx.jk(D, chunk_bytes=80 << 30)
elif mode == "checksum":
    ctx.set_checksums(True)
    st = ctx.all_unique(chunk_bytes=80 << 30)
elif mode == "wide":
    from libcint_b200.basis import class_sweep_basis
    a2, b2, e2 = class_sweep_basis(lmax=5, nctr=1)
    c2 = cb.Context(a2, b2, e2)
    q = np.tile(np.array([[3, 6 + 3, 12 + 3, 18 + 3]], np.int32), (4000, 1))      # (ff|ff)
    c2.int2e_batch(q)
    q = np.tile(np.array([[4, 6 + 4, 12 + 4, 18 + 4]], np.int32), (800, 1))       # (gg|gg)
    c2.int2e_batch(q)
    st = np.zeros(16)
else:
    st = ctx.all_unique(chunk_bytes=80 << 30)
    np.save(os.path.join("gpurun_out", "launch_rows.npy"), ctx.launch_rows())
print("gpu ms", st[7], "launches", st[4])
PY
for what in "$@"; do
case $what in
launches)
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$ROUND.csv python /tmp/onejob.py jk > gpurun_out/launches_$ROUND.log 2>&1
  ;;
pipes)
  ncu --metrics gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed \
      --clock-control none --csv --log-file gpurun_out/pipes_$ROUND.csv python /tmp/onejob.py plain > gpurun_out/pipes_$ROUND.log 2>&1
  ;;
full)
  capture() {   # name, mangled-regex, skip, mode
    ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$2 -s $3 -c 1 -f -o /tmp/prof_$1 python /tmp/onejob.py $4 > gpurun_out/prof_$1_$ROUND.log 2>&1
    ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/prof_$1_${ROUND}_raw.csv 2>/dev/null
    ncu -i /tmp/prof_$1.ncu-rep --page details > gpurun_out/prof_$1_${ROUND}_details.txt 2>/dev/null
  }
  capture reg_psps eri_reg_kernelILi1ELi0ELi1ELi0ELi2ELi2E 4 plain
  capture coop_dpdp eri_coop_kernelILi2ELi1ELi2ELi1E 4 plain
  capture coop_dsdp eri_coop_kernelILi2ELi1ELi2ELi0E 4 plain
  capture coop_dddd eri_coop_kernelILi2ELi2ELi2ELi2E 4 plain
  capture jk_rows3 jk_rows_kernelILi3E 3 jk
  capture jk_cols jk_cols_kernel 3 jk
  capture rowsum tile_rowsum_kernel 3 checksum
  capture wide_ffff eri_wide_kernelILi3ELi3E 0 wide
  ;;
esac
done
ls -la gpurun_out | tail -30
