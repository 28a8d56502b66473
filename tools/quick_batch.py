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
    # the C entry point alone (no numpy size computation of the Python mirror): packed output, device buffer, no flags
    import ctypes
    f = ctx.lib.cintb200_int2e_batch
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        rc = f(ctx.handle, 0, q.ctypes.data_as(ctypes.c_void_p), nq, None, ctypes.c_void_p(buf.data_ptr()), 1, None)
        torch.cuda.synchronize(); dt = time.time() - t0
    print("   C call only: %.1f ms -> %.3g quartets/s, %.3g integrals/s (rc %d)" % (dt * 1e3, nq / dt, tot / dt, rc))
# structured list: every bra of a set with every ket of a set (what screened direct-SCF lists look like), same API
import ctypes
nb = len(bas)
pairs = np.array([(i, j) for i in range(nb) for j in range(i + 1)], np.int32)
for nbra, nket in ((2000, 500), (20000, 100)):
    bra = pairs[rng.choice(len(pairs), nbra, replace=False)]
    ket = pairs[rng.choice(len(pairs), nket, replace=False)]
    q = np.concatenate([np.repeat(bra, nket, axis=0), np.tile(ket, (nbra, 1))], axis=1).astype(np.int32)
    perm = rng.permutation(len(q))
    q = np.ascontiguousarray(q[perm])                      # order of the list does not matter
    tot = int(np.prod(dim[q], axis=1).sum())
    buf = torch.empty(tot, dtype=torch.float64, device="cuda")
    for it in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        ctx.int2e_batch(q, device_ptr=buf.data_ptr())
        torch.cuda.synchronize(); dt = time.time() - t0
    print("structured list %d bras x %d kets: %d quartets, %.3g integrals in %.1f ms -> %.3g quartets/s, %.3g integrals/s" % (
        nbra, nket, len(q), tot, dt * 1e3, len(q) / dt, tot / dt))
