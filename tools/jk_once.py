#!/usr/bin/env python
"""One J/K digestion pass of the C60 job (for ncu launch lists of the consumer kernels / host-phase timing)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import libcint_b200 as cb
chunk = int(float(sys.argv[1]) * (1 << 30)) if len(sys.argv) > 1 else 80 << 30
atm, bas, env = cb.load_fixture("c60_ccpvdz")
_, _, D, _ = cb.job_weights(840)
t0 = time.perf_counter()
ctx = cb.Context(atm, bas, env)
t1 = time.perf_counter()
vj, vk, st = ctx.jk(D, chunk_bytes=chunk)
t2 = time.perf_counter()
vj, vk, st2 = ctx.jk(D, chunk_bytes=chunk)
t3 = time.perf_counter()
ctx.close()
t4 = time.perf_counter()
print(json.dumps({"chunk_gb": chunk / 2**30, "create_s": t1 - t0, "first_jk_s": t2 - t1, "second_jk_s": t3 - t2, "close_s": t4 - t3,
                  "gpu_ms_first": float(st[7]), "gpu_ms_second": float(st2[7])}))
