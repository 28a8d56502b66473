#!/usr/bin/env python
"""Golden whole-job fingerprints from the UNMODIFIED reference (oracle/_ref/libcint_ref.so via oracle/_ref/ref_golden):
per-bra-pair sums S/A/F over every quartet of the reference benchmark loop, and J/K for the formula density of
oracle/ref_golden.c, stored as projections on a fixed probe matrix (small fixture).

  python tools/make_golden_job.py c60_ccpvdz        -> tests/golden/job_c60_ccpvdz.npz   (~15 min on 8 cores)
  python tools/make_golden_job.py c2h6_ccpvdz ...   -> small molecules, seconds

Definitions shared with the tests: job_weights() below.
"""
import os
import subprocess
import sys
import tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def job_weights(nao, maxrb):
    """h[r], g[c,d], D[a,b], probe U[nao,8] -- the formulas of oracle/ref_golden.c."""
    r = np.arange(maxrb)
    a = np.arange(nao)
    h = np.cos(0.91 * r + 0.3)
    g = np.cos(0.37 * a[:, None] + 0.61 * a[None, :] + 0.5)
    D = np.cos(0.37 * (a[:, None] + a[None, :]) + 0.2) + 0.5 * np.cos(0.11 * (a[:, None] - a[None, :]))
    U = np.cos(0.13 * (a[:, None] + 1) * (np.arange(8)[None, :] + 1) + 0.7)
    return h, g, D, U


def run(name, stride=1, phase=0):
    import libcint_b200 as cb
    from bench import dump_basis_bin
    atm, bas, env = cb.load_fixture(name)
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_golden")
    lib = os.path.join(ROOT, "oracle", "_ref", "libcint_ref.so")
    with tempfile.TemporaryDirectory() as tmp:
        bb, ob = os.path.join(tmp, "basis.bin"), os.path.join(tmp, "out.bin")
        dump_basis_bin(bb, atm, bas, env)
        out = subprocess.run([exe, lib, bb, ob, str(stride), str(phase)], capture_output=True, text=True, check=True)
        print(out.stdout.strip())
        raw = np.fromfile(ob, dtype=np.uint8)
    npair, nao = np.frombuffer(raw[:16], dtype=np.int64)
    v = np.frombuffer(raw[16:], dtype=np.float64)
    S, A, F = v[:npair], v[npair:2 * npair], v[2 * npair:3 * npair]
    J = v[3 * npair:3 * npair + nao * nao].reshape(nao, nao)
    K = v[3 * npair + nao * nao:].reshape(nao, nao)
    return dict(S=S.copy(), A=A.copy(), F=F.copy(), J=J.copy(), K=K.copy())


if __name__ == "__main__":
    for name in sys.argv[1:]:
        res = run(name)
        nao = res["J"].shape[0]
        _, _, D, U = job_weights(nao, 1)
        out = os.path.join(ROOT, "tests", "golden", "job_%s.npz" % name)
        np.savez_compressed(out, S=res["S"], A=res["A"], F=res["F"], JU=res["J"] @ U, KU=res["K"] @ U,
                            trJD=float(np.sum(res["J"] * D)), trKD=float(np.sum(res["K"] * D)),
                            Jdiag=np.diag(res["J"]).copy(), Kdiag=np.diag(res["K"]).copy())
        print("wrote", out, os.path.getsize(out), "bytes")
