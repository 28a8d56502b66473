"""Time the full density-fitting job (config 3 stand-in: C60, cc-pVTZ orbital shells, s..g auxiliary shells) on one GPU."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import libcint_b200 as cb
from libcint_b200.basis import c60_df_basis

atm, bas, env, norb = c60_df_basis()
ctx = cb.Context(atm, bas, env)
gen = len(sys.argv) > 1 and sys.argv[1] == "generic"
if gen:
    ctx.force_generic(True)
for it in range(3 if not gen else 1):
    t0 = time.time()
    st = ctx.int3c2e_all(norb, chunk_bytes=80 << 30)
    print("pass %d: gpu %.1f ms wall %.1f ms triples %.4g integrals %.4g prim %.4g model %.3g TFLOP/s store %.0f GB/s launches %d reg/coop %d"
          % (it, st[7], (time.time() - t0) * 1e3, st[0], st[1], st[2], st[6] / st[7] / 1e9, st[1] * 8 / st[7] / 1e6, st[4], st[8]))
st, rows = (None, None)
ctx.lib.cintb200_debug_profile(ctx.handle, 1)
st = ctx.int3c2e_all(norb, chunk_bytes=80 << 30)
ctx.lib.cintb200_debug_profile(ctx.handle, 0)
n = ctx.lib.cintb200_debug_profile_rows(ctx.handle, None, 0)
rows = np.zeros((n, 12))
ctx.lib.cintb200_debug_profile_rows(ctx.handle, rows.ctypes.data_as(__import__('ctypes').c_void_p), n)
rows = rows[np.argsort(-rows[:, 7])]
print("serialised total %.1f ms" % rows[:, 7].sum())
for r in rows[:40]:
    print("(%d%d|%d%d) nct %d ncu %d kind %d  %.2f ms  triples %.3g  GFLOP/s %.0f  GB/s %.0f" % (
        r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8], r[10] / r[7] / 1e6,
        r[8] * (2 * r[0] + 1) * (2 * r[1] + 1) * (2 * r[2] + 1) * r[4] * r[5] * 8 / r[7] / 1e6))
