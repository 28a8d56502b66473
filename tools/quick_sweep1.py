#!/usr/bin/env python
"""One class of the class-sweep bench: wall clock per call and the library's host phases (CINTB200_TIMING=1).
usage: quick_sweep1.py la lb lc ld nctr [reps]"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import libcint_b200 as cb
from libcint_b200.basis import class_sweep_basis
cls = [int(x) for x in sys.argv[1:5]]
nctr = int(sys.argv[5])
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 20000
atm, bas, env = class_sweep_basis(lmax=5, nctr=nctr)
c = cb.Context(atm, bas, env)
sh = [cen * 6 + l for cen, l in enumerate(cls)]
q = np.tile(np.array(sh, np.int32), (reps, 1))
n1 = int(np.prod([(2 * l + 1) * nctr for l in cls]))
dbuf = torch.empty(n1 * reps, dtype=torch.float64, device="cuda")
ts = []
for k in range(4):
    print("---- call", k, file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    c.int2e_batch(q, device_ptr=dbuf.data_ptr())
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t0)
print(json.dumps({"class": cls, "nctr": nctr, "reps": reps, "ms_per_call": [1e3 * t for t in ts], "us_per_quartet": 1e6 * min(ts[1:]) / reps}))
