#!/bin/bash
# GPU box: compare kernel build variants (CINTB200_LIB=...) on the C60 whole job.
mkdir -p gpurun_out
for v in "" _mb3 _mb4; do
  lib=$PWD/libcint_b200/libcint_b200$v.so
  [ -f $lib ] || continue
  echo "=== variant '$v'"
  CINTB200_LIB=$lib python tools/profile_c60.py > gpurun_out/profile_c60$v.txt 2>&1
  head -1 gpurun_out/profile_c60$v.txt
  CINTB200_LIB=$lib python - <<'PY'
import sys, time
sys.path.insert(0, '.')
import libcint_b200 as cb
atm, bas, env = cb.load_fixture("c60_ccpvdz")
ctx = cb.Context(atm, bas, env)
for gb in (80,):
    ctx.all_unique(chunk_bytes=gb << 30)
    st = ctx.all_unique(chunk_bytes=gb << 30)
    print("multi-stream chunk %d GB: gpu %.1f ms launches %d chunks %d" % (gb, st[7], st[4], st[9]))
PY
done
