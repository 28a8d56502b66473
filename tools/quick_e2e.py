"""End-to-end C60 pass only (host arrays in, every tile copied to a pinned host buffer): seconds per pass."""
import sys, time
sys.path.insert(0, '.')
import torch
import libcint_b200 as cb
atm, bas, env = cb.load_fixture("c60_ccpvdz")
chunk = 16 << 30
sink = torch.empty(chunk // 8, dtype=torch.float64, pin_memory=True)
def step():
    c = cb.Context(atm, bas, env)
    st = c.all_unique(chunk_bytes=chunk, host_sink=sink.data_ptr())
    c.close()
    return st
step()
torch.cuda.synchronize(); t0 = time.time()
st = step()
torch.cuda.synchronize(); dt = time.time() - t0
print("e2e %.3f s per pass, D2H %.1f GB/s" % (dt, st[5] / dt / 1e9))
