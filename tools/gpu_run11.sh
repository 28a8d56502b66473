set -x
cd $GRAFT_REPO_ROOT
for mg in 1 0; do CINTB200_MERGE=$mg timeout 200 python tools/time_variant.py 80; done > gpurun_out/r2k_time.log 2>&1; cat gpurun_out/r2k_time.log
CINTB200_MERGE=1 timeout 200 python tools/time_variant.py 40 >> gpurun_out/r2k_time.log 2>&1; tail -1 gpurun_out/r2k_time.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2k_tests.log 2>&1; tail -5 gpurun_out/r2k_tests.log
CINTB200_TIMING=1 timeout 100 python - > gpurun_out/r2k_list_timing.log 2>&1 <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, torch, ctypes, time
import libcint_b200 as cb
atm, bas, env = cb.load_fixture("c60_ccpvdz")
ctx = cb.Context(atm, bas, env)
q = np.tile(np.array([[3, 2, 1, 0]], np.int32), (20000, 1))
buf = torch.empty(20000 * 200, dtype=torch.float64, device="cuda")
f = ctx.lib.cintb200_int2e_batch
for it in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    rc = f(ctx.handle, 0, q.ctypes.data_as(ctypes.c_void_p), len(q), None, ctypes.c_void_p(buf.data_ptr()), 1, None)
    torch.cuda.synchronize(); print("call %d: %.2f ms rc %d" % (it, 1e3 * (time.time() - t0), rc), file=sys.stderr)
PY
tail -20 gpurun_out/r2k_list_timing.log
