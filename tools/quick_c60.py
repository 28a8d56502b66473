import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import libcint_b200 as cb
atm, bas, env = cb.load_fixture("c60_ccpvdz")
t=time.time(); ctx = cb.Context(atm, bas, env); print('create', time.time()-t)
for it in range(3):
    t=time.time(); st = ctx.all_unique(chunk_bytes=16<<30); dt=time.time()-t
    print('wall %.3f s gpu %.1f ms quartets %.4g integrals %.5g prim %.4g launches %d (reg %d) chunks %d flops %.3e -> %.3e int/s, %.2f TFLOP/s model' % (dt, st[7], st[0], st[1], st[2], st[4], st[8], st[9], st[6], st[1]/(st[7]*1e-3), st[6]/(st[7]*1e-3)/1e12))
