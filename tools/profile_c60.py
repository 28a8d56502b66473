"""Per-class time/FLOP table of the C60 whole job (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import libcint_b200 as cb
name = sys.argv[1] if len(sys.argv) > 1 else "c60_ccpvdz"
atm, bas, env = cb.load_fixture(name)
CH = int(os.environ.get("CHUNK_GB", "16")) << 30
ctx = cb.Context(atm, bas, env)
ctx.all_unique(chunk_bytes=CH)
st, rows = ctx.profile(chunk_bytes=CH)
tot = rows[:, 7].sum()
print("total %.1f ms (events sum %.1f) model %.3e flop" % (st[7], tot, st[6]))
print("class nct ncu kind      ms    %%   quartets   primq   GFLOP/s(model)  ns/primq  launches")
for r in sorted(rows, key=lambda r: -r[7]):
    print("(%d%d|%d%d) %d %d %s %9.2f %5.1f %9.3g %9.3g %9.1f %9.3f %5d" % (r[0], r[1], r[2], r[3], r[4], r[5], {0: "GEN", 1: "reg", 2: "coop"}[int(r[6])],
          r[7], 100 * r[7] / tot, r[8], r[9], r[10] / r[7] / 1e6, r[7] * 1e6 / max(r[9], 1), r[11]))
