#!/usr/bin/env python
"""Generate tests/golden/*.npz from the compiled reference (oracle/_ref) -- run in the build container.

  c60_blocks.npz     full blocks of int2e_sph for the SURVEY Appendix-D quartets + 120 seeded random
                     quartets of C60/cc-pVDZ (examples/time_c60.c basis)
  testbasis.npz      per-quartet checksums (sum|v|, sum v*cos(n)) of int2e_sph over ALL 8^4 quartets of the
                     reference test basis (testsuite/test_cint.py:46-137), int3c2e_sph over all 8^3 triples,
                     int2e_cart for a subset, and the LR (omega=0.5) variant
  derivs.npz         checksums of int2c2e_sph, int2c2e_ip1/ip2_sph, int3c2e_ip1/ip2_sph, int2e_ip1_sph/_cart over the test basis and
                     a few full blocks
  rys_mpmath.npz     Rys roots/weights from the reference's 100-digit mpmath implementation
                     (scripts/rys_roots.py:197) for nroots 1..11 on a grid of x
"""
import os, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_util as ou
from libcint_b200 import load_fixture
from libcint_b200.basis import reference_test_basis

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
assert ou.ref() is not None, "build oracle/_ref first (make -C oracle ref)"


def fp(v):
    return np.array([np.abs(v).sum(), (v * np.cos(np.arange(v.size))).sum()])


def c60():
    atm, bas, env = load_fixture("c60_ccpvdz")
    shls = [(0, 0, 0, 0), (1, 0, 1, 0), (2, 2, 2, 2), (4, 4, 4, 4), (4, 3, 2, 0), (299, 150, 77, 3),
            (149, 148, 4, 2), (5, 0, 295, 290)]
    rng = np.random.default_rng(60)
    for _ in range(120):
        shls.append(tuple(int(x) for x in rng.integers(0, 300, 4)))
    vals = ou.eval_many("ref", "int2e_sph", shls, atm, bas, env)
    np.savez_compressed(os.path.join(OUT, "c60_blocks.npz"), shls=np.array(shls, np.int32),
                        offsets=np.cumsum([0] + [v.size for v in vals]), values=np.concatenate(vals))


def testbasis():
    atm, bas, env = reference_test_basis(with_fit_shells=True)
    q = [(i, j, k, l) for i in range(8) for j in range(8) for k in range(8) for l in range(8)]
    f4 = np.array([fp(v) for v in ou.eval_many("ref", "int2e_sph", q, atm, bas, env)])
    t = [(i, j, k) for i in range(8) for j in range(8) for k in range(8)]
    f3 = np.array([fp(v) for v in ou.eval_many("ref", "int3c2e_sph", t, atm, bas, env)])
    qc = q[::7]
    fc = np.array([fp(v) for v in ou.eval_many("ref", "int2e_cart", qc, atm, bas, env)])
    env_lr = env.copy()
    env_lr[8] = 0.5
    ql = q[::5]
    fl = np.array([fp(v) for v in ou.eval_many("ref", "int2e_sph", ql, atm, bas, env_lr)])
    np.savez_compressed(os.path.join(OUT, "testbasis.npz"), q4=np.array(q, np.int32), f4=f4,
                        q3=np.array(t, np.int32), f3=f3, qcart=np.array(qc, np.int32), fcart=fc,
                        qlr=np.array(ql, np.int32), flr=fl, omega_lr=0.5)


def derivs():
    """derivs.npz: per-tuple checksums of the 2-centre metric and of the first-derivative integrals over the reference
    test basis (all 8^2 / 8^3 tuples, every 9th of the 8^4 quartets) + full blocks of a few tuples"""
    import itertools
    atm, bas, env = reference_test_basis(with_fit_shells=True)
    out = {}
    for name, nc, step in (("int2c2e_sph", 2, 1), ("int2c2e_ip1_sph", 2, 1), ("int2c2e_ip2_sph", 2, 1), ("int3c2e_ip1_sph", 3, 1),
                           ("int3c2e_ip2_sph", 3, 1), ("int2e_ip1_sph", 4, 9), ("int2e_ip1_cart", 4, 31)):
        t = list(itertools.product(range(8), repeat=nc))[::step]
        out["q_" + name] = np.array(t, np.int32)
        out["f_" + name] = np.array([fp(v) for v in ou.eval_many("ref", name, t, atm, bas, env)])
    full = [("int2e_ip1_sph", (1, 2, 5, 3)), ("int2e_ip1_sph", (3, 0, 2, 7)), ("int3c2e_ip2_sph", (2, 5, 3)), ("int2c2e_sph", (3, 2))]
    for n, (name, sh) in enumerate(full):
        out["full%d_name" % n] = name
        out["full%d_shls" % n] = np.array(sh, np.int32)
        out["full%d_vals" % n] = ou.eval_tuple("ref", name, sh, atm, bas, env)[0]
    out["nfull"] = len(full)
    np.savez_compressed(os.path.join(OUT, "derivs.npz"), **out)


def rys():
    sys.path.insert(0, "/root/reference/scripts")
    import rys_roots as rr
    import mpmath
    mpmath.mp.dps = 60
    xs = [0.0, 1e-8, 1e-3, 0.05, 0.7, 2.9, 7.7, 14.2, 19.9, 33.3, 39.99, 40.01, 61.3, 89.9, 90.1, 133.0, 777.0]
    rows = []
    for n in range(1, 12):
        for x in xs:
            r, w = rr.rys_roots_weights(n, x)
            t2 = [float(v / (1 + v)) for v in r]     # the script returns u = t^2/(1-t^2)
            rows.append([n, x] + t2 + [0.0] * (11 - n) + [float(v) for v in w] + [0.0] * (11 - n))
    np.savez_compressed(os.path.join(OUT, "rys_mpmath.npz"), table=np.array(rows))


if __name__ == "__main__":
    which = sys.argv[1:] or ["c60", "testbasis", "derivs", "rys"]
    for w in which:
        globals()[w]()
        print("wrote", w)
