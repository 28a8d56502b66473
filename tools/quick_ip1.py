#!/usr/bin/env python
"""Wall clock and host phases (CINTB200_TIMING=1) of the int2e_ip1 gradient loop of bench.py on C2H6 cc-pVQZ."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import libcint_b200 as cb
atm, bas, env = cb.load_fixture(sys.argv[1] if len(sys.argv) > 1 else "c2h6_ccpvqz")
c = cb.Context(atm, bas, env)
nb = len(bas)
dims = np.array([(2 * int(b[1]) + 1) * int(b[3]) for b in bas])
nao = int(dims.sum())
dbuf = torch.empty(3 * nao * nao * int(dims.max()) * nao, dtype=torch.float64, device="cuda")
def one_pass(verbose=False):
    ms, t0 = 0.0, time.perf_counter()
    for k in range(nb):
        if verbose: print("---- ket shell", k, "l", int(bas[k][1]), flush=True); sys.stderr.flush()
        _, st = c.ip1_block((0, nb, 0, nb, k, k + 1, 0, k + 1), device_ptr=dbuf.data_ptr())
        ms += float(st[7])
    torch.cuda.synchronize()
    return time.perf_counter() - t0, ms
one_pass()
os.environ["CINTB200_TIMING"] = "0"
w, ms = one_pass()
print(json.dumps({"wall_s": w, "eri_kernel_ms": ms, "integrals_per_s": float(nao) ** 4 / 2 * 3 / w}))
