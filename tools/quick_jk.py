#!/usr/bin/env python
"""Timing of the tile consumers on the C60 job: plain pass, pass with checksums, J only, J + K (device-resident dm / vj / vk)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import libcint_b200 as cb

chunk = int(float(sys.argv[1]) * (1 << 30)) if len(sys.argv) > 1 else 80 << 30
atm, bas, env = cb.load_fixture("c60_ccpvdz")
_, _, D, _ = cb.job_weights(840)
ctx = cb.Context(atm, bas, env)
for _ in range(2):
    st = ctx.all_unique(chunk_bytes=chunk)
out = {"plain_ms": float(st[7])}
ctx.set_checksums(True)
st = ctx.all_unique(chunk_bytes=chunk); st = ctx.all_unique(chunk_bytes=chunk)
out["checksums_ms"] = float(st[7])
ctx.set_checksums(False)
d = torch.tensor(D, device="cuda"); vj = torch.zeros_like(d); vk = torch.zeros_like(d)
for with_k in (False, True):
    for _ in range(2):
        t0 = time.perf_counter()
        _, _, st = ctx.jk(rank=0, nranks=1, chunk_bytes=chunk, with_k=with_k, device_ptrs=(d.data_ptr(), vj.data_ptr(), vk.data_ptr()))
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
    out["jk_ms" if with_k else "j_ms"] = float(st[7])
    out["jk_wall_ms" if with_k else "j_wall_ms"] = 1e3 * wall
# host-in / host-out call incl. context build (what bench.py's e2e times)
t0 = time.perf_counter()
c2 = cb.Context(atm, bas, env)
vjh, vkh, st = c2.jk(D, chunk_bytes=chunk)
out["e2e_jk_s"] = time.perf_counter() - t0
c2.close()
print(json.dumps(out))
