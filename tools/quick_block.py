"""Throughput of the dense shell-slice block call vs the list-mode batch call on C60 (device-resident output)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import libcint_b200 as cb
atm, bas, env = cb.load_fixture("c60_ccpvdz")
ctx = cb.Context(atm, bas, env)
ao = np.concatenate([[0], np.cumsum([(2 * int(b[1]) + 1) * int(b[3]) for b in bas])])
sl = (0, 60, 0, 60, 0, 300, 0, 300)
n = int((ao[60]) ** 2 * 840 * 840)
buf = torch.empty(n, dtype=torch.float64, device="cuda")
for it in range(3):
    t0 = time.time()
    _, st = ctx.int2e_block(sl, device_ptr=buf.data_ptr())
    torch.cuda.synchronize()
    print("block %s: wall %.1f ms gpu %.1f ms quartets %.3g integrals %.3g -> %.3g integrals/s (gpu), %.2f TFLOP/s model" % (
        sl, (time.time() - t0) * 1e3, st[7], st[0], st[1], st[1] / st[7] * 1e3, st[6] / st[7] / 1e9))
rng = np.random.default_rng(0)
q = rng.integers(0, 300, (200000, 4)).astype(np.int32)
sizes = np.array([np.prod(cb.shell_dims(bas, s)) for s in q])
buf2 = torch.empty(int(sizes.sum()), dtype=torch.float64, device="cuda")
for it in range(2):
    t0 = time.time()
    ctx.int2e_batch(q, device_ptr=buf2.data_ptr())
    torch.cuda.synchronize()
    dt = time.time() - t0
    print("list-mode batch: %d quartets, %.3g integrals in %.1f ms -> %.3g integrals/s" % (len(q), sizes.sum(), dt * 1e3, sizes.sum() / dt))
