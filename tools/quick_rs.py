"""Whole C60 job with the range-separated operator (long range erf, short range erfc) on the specialised kernels."""
import sys
sys.path.insert(0, '.')
import numpy as np
import libcint_b200 as cb
atm, bas, env = cb.load_fixture("c60_ccpvdz")
for omega in (0.0, 0.3, -0.3):
    e = env.copy(); e[8] = omega
    ctx = cb.Context(atm, bas, e)
    for it in range(3):
        st = ctx.all_unique(chunk_bytes=80 << 30)
    kinds = sorted(set(int(r[7]) for r in ctx.launch_rows()))
    print("omega %+.1f: %.1f ms per pass, %.3g integrals/s, launch kinds %s" % (omega, st[7], 6.2234e10 / st[7] * 1e3, kinds))
    ctx.close()
