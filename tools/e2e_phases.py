#!/usr/bin/env python
"""Wall-clock phases of the end-to-end J/K step of bench.py (context build, density H2D, jk, D2H, teardown), 4 steps."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import libcint_b200 as cb
atm, bas, env = cb.load_fixture("c60_ccpvdz")
nao = 840
_, _, Dm, _ = cb.job_weights(nao)
dm_host = torch.from_numpy(np.ascontiguousarray(Dm)).pin_memory()
jk_host = torch.empty((2, nao, nao), dtype=torch.float64).pin_memory()
chunk = 80 << 30
for step in range(4):
    t = [time.perf_counter()]
    c2 = cb.Context(atm, bas, env, device=0); t.append(time.perf_counter())
    d_dm = dm_host.to("cuda", non_blocking=True)
    d_jk = torch.empty((2, nao, nao), dtype=torch.float64, device="cuda")
    torch.cuda.current_stream().synchronize(); t.append(time.perf_counter())
    _, _, s2 = c2.jk(rank=0, nranks=1, chunk_bytes=chunk, device_ptrs=(d_dm.data_ptr(), d_jk[0].data_ptr(), d_jk[1].data_ptr())); t.append(time.perf_counter())
    jk_host.copy_(d_jk); torch.cuda.synchronize(); t.append(time.perf_counter())
    c2.close(); t.append(time.perf_counter())
    d = np.diff(t)
    print(json.dumps({"step": step, "context_s": d[0], "h2d_s": d[1], "jk_s": d[2], "d2h_s": d[3], "close_s": d[4], "total_s": t[-1] - t[0], "gpu_ms": float(s2[7])}), flush=True)
