"""Throughput of the list-mode batch call (generic kernel) on C60, device-resident output: random quartets."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import libcint_b200 as cb
atm, bas, env = cb.load_fixture("c60_ccpvdz")
ctx = cb.Context(atm, bas, env)
rng = np.random.default_rng(0)
for nq in (20000, 200000, 1000000):
    q = rng.integers(0, 300, (nq, 4)).astype(np.int32)
    dim = (2 * bas[:, 1] + 1) * bas[:, 3]
    tot = int(np.prod(dim[q], axis=1).sum())
    buf = torch.empty(tot, dtype=torch.float64, device="cuda")
    for it in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        ctx.int2e_batch(q, device_ptr=buf.data_ptr())
        torch.cuda.synchronize(); dt = time.time() - t0
    print("list-mode batch: %d quartets, %.3g integrals in %.1f ms -> %.3g quartets/s, %.3g integrals/s" % (nq, tot, dt * 1e3, nq / dt, tot / dt))
