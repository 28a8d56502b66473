#!/usr/bin/env python
"""Small jobs that touch every kernel family, for compute-sanitizer (memcheck / racecheck) runs on the GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import libcint_b200 as cb
from libcint_b200.basis import class_sweep_basis
atm, bas, env = cb.load_fixture("c2h6_ccpvdz")
ctx = cb.Context(atm, bas, env)
ctx.set_checksums(True)
st = ctx.all_unique(chunk_bytes=300_000)                    # register + cooperative kernels, merged launches, checksum consumer
S, A, F = ctx.job_checksums()
ctx.set_checksums(False)
_, _, D, _ = cb.job_weights(58)
vj, vk, _ = ctx.jk(D, chunk_bytes=300_000)                  # J/K digestion kernels
ctx.all_unique(cart=True)                                   # Cartesian instantiations
rng = np.random.default_rng(0)
q = rng.integers(0, len(bas), size=(5000, 4)).astype(np.int32)
v, o, s, nz = ctx.int2e_batch(q)                            # device-side list bookkeeping
blk, _ = ctx.int2e_block((0, 10, 0, 10, 5, 20, 0, 8))
g, _ = ctx.ip1_block((0, 6, 0, 8, 3, 12, 0, 5))             # derivative blocks
a2, b2, e2 = class_sweep_basis(lmax=5, nctr=1)
c2 = cb.Context(a2, b2, e2)
for sh in ([3, 6 + 3, 12 + 3, 18 + 3], [4, 6 + 4, 12 + 2, 18 + 1], [5, 6 + 0, 12 + 4, 18 + 4]):      # wide kernel + catch-all epilogue
    c2.int2e_batch(np.tile(np.array(sh, np.int32), (40, 1)))
a3, b3, e3 = class_sweep_basis(lmax=3, nctr=2)             # catch-all kernel with general contractions: primitive batches, split contraction updates
c3 = cb.Context(a3, b3, e3)
for sh in ([1, 4 + 1, 8 + 0, 12 + 0], [2, 4 + 1, 8 + 1, 12 + 0], [3, 4 + 2, 8 + 2, 12 + 1]):
    c3.int2e_batch(np.tile(np.array(sh, np.int32), (24, 1)))
e3s = e3.copy(); e3s[8] = -0.4                              # short-range operator: the 2N-point rule in the batched phases
c4 = cb.Context(a3, b3, e3s)
c4.int2e_batch(np.tile(np.array([2, 4 + 2, 8 + 1, 12 + 1], np.int32), (24, 1)))
print("sanitize job done: launches", int(st[4]), "checksum", float(S.sum()), "trJ", float(np.trace(vj)))
