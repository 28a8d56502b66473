"""Synthetic basis builders shared by tests and bench (inputs only -- no integral code here)."""
import math
import numpy as np

PTR_ENV_START = 20


def gto_norm(l, a):
    """Radial normalisation of r^l exp(-a r^2) -- closed form of the reference's CINTgto_norm (src/misc.c:86)."""
    p = l + 1.5
    return 1.0 / math.sqrt(math.gamma(p) / (2.0 * (2.0 * a) ** p))


def reference_test_basis(with_fit_shells=False):
    """The 4-atom p/d/f/g + s/p/d/f basis of the reference's known-answer tests
    (testsuite/test_cint.py:46-137; testsuite/test_3c2e.py:133-159 adds two zero-exponent s shells)."""
    natm = 4
    atm = np.zeros((natm, 6), np.int32)
    bas = np.zeros((10 if with_fit_shells else 8, 8), np.int32)
    env = np.zeros(120)
    off = PTR_ENV_START
    for i in range(natm):
        atm[i, 0] = (i + 1) * 2
        atm[i, 1] = off
        env[off:off + 3] = [.2 * (i + 1), .3 + (i + 1) * .5, .1 - (i + 1) * .5]
        off += 3
    off0 = off
    spec = [(0, 1, [1.], [1.]), (1, 2, [5., 3.], [1., 2., 4., 1.]), (2, 3, [1.], [1.]), (3, 4, [.5], [1.])]
    for nh, (ia, l, exps, coefs) in enumerate(spec):
        bas[nh, 0] = ia
        bas[nh, 1] = l
        bas[nh, 4] = l
        bas[nh, 2] = len(exps)
        bas[nh, 3] = len(coefs) // len(exps)
        bas[nh, 5] = off
        env[off:off + len(exps)] = exps
        off += len(exps)
        bas[nh, 6] = off
        env[off:off + len(coefs)] = coefs
        off += len(coefs)
    nh = 4
    n = off - off0
    env[off:off + n] = env[off0:off0 + n]
    for i in range(nh):
        bas[i + nh] = bas[i]
        bas[i + nh, 1] -= 1
        bas[i + nh, 4] = -bas[i, 4]
        bas[i + nh, 5] += n
        bas[i + nh, 6] += n
        env[bas[i + nh, 6]] /= 2 * env[bas[i, 5]]
    e0, e1 = env[bas[1, 5]], env[bas[1, 5] + 1]
    c = env[bas[1, 6]:bas[1, 6] + 4].copy()
    env[bas[5, 6]:bas[5, 6] + 4] = [c[0] / (2 * e0), c[1] / (2 * e1), c[2] / (2 * e0), c[3] / (2 * e1)]
    off += n
    if with_fit_shells:
        for k in range(2):
            r = 8 + k
            bas[r, 0] = 0
            bas[r, 2] = bas[r, 3] = 1
            bas[r, 5] = off
            env[off] = 0.0
            off += 1
            bas[r, 6] = off
            env[off] = 2 * math.sqrt(math.pi)
            off += 1
    return atm, bas, env[:off + 1].copy()


def class_sweep_basis(lmax=4, nprim=3, nctr=2, seed=20241017):
    """Config 4: four centres of the reference test geometry, one shell per (centre, l),
    `nprim` primitives x `nctr` contractions, exponents {2.0,0.8,0.3}(1+0.1 l), seeded coefficients."""
    rng = np.random.default_rng(seed)
    natm = 4
    atm = np.zeros((natm, 6), np.int32)
    env = [0.0] * PTR_ENV_START
    for i in range(natm):
        atm[i, 0] = 1
        atm[i, 1] = len(env)
        env += [.2 * (i + 1), .3 + (i + 1) * .5, .1 - (i + 1) * .5]
    bas = []
    base = [2.0, 0.8, 0.3, 0.11, 5.1][:nprim]
    for ia in range(natm):
        for l in range(lmax + 1):
            exps = [e * (1 + 0.1 * l) for e in base]
            pe = len(env)
            env += exps
            pc = len(env)
            for c in range(nctr):
                env += [rng.uniform(0.2, 1.0) * gto_norm(l, a) for a in exps]
            bas.append([ia, l, nprim, nctr, 0, pe, pc, 0])
    return atm, np.array(bas, np.int32), np.array(env)


def unique_quartets(nbas):
    """Shell quartets of the reference benchmark loop (examples/time_c60.c:200-215): i>=j, k>=l, k<=i."""
    out = []
    for i in range(nbas):
        for j in range(i + 1):
            for k in range(i + 1):
                for l in range(k + 1):
                    out.append((i, j, k, l))
    return np.array(out, np.int32)


def c60_df_basis(max_atoms=60, aux_lmax=4):
    """Config 3 stand-in (SURVEY 8d / hard part 7): C60 geometry of examples/time_c60.c with the carbon cc-pVTZ
    shells of examples/time_c2h6.c (taken from the committed fixtures) as orbital basis, plus an EVEN-TEMPERED
    auxiliary basis per atom (the def2-universal-JKFIT exponents are not available offline):
    l = 0..aux_lmax, exponents 0.25 * 2.2^k, k = 0..(5 - l), one uncontracted primitive per shell.
    Returns atm, bas, env, n_orbital_shells."""
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
    c60 = np.load(os.path.join(here, "c60_ccpvdz.npz"))
    tz = np.load(os.path.join(here, "c2h6_ccpvtz.npz"))
    atm60, env60 = c60["atm"], c60["env"]
    tbas, tenv = tz["bas"], tz["env"]
    cshells = [b for b in tbas if b[0] == 0]              # shells of the first carbon atom
    natm = min(max_atoms, 60)
    atm = np.zeros((natm, 6), np.int32)
    env = [0.0] * PTR_ENV_START
    for i in range(natm):
        atm[i, 0] = 6
        atm[i, 1] = len(env)
        env += list(env60[atm60[i, 1]:atm60[i, 1] + 3])
    bas = []
    ptr = {}
    for b in cshells:                                     # share exponent/coefficient storage between atoms
        key = (int(b[5]), int(b[6]))
        if key not in ptr:
            pe = len(env)
            env += list(tenv[b[5]:b[5] + b[2]])
            pc = len(env)
            env += list(tenv[b[6]:b[6] + b[2] * b[3]])
            ptr[key] = (pe, pc)
    for i in range(natm):
        for b in cshells:
            pe, pc = ptr[(int(b[5]), int(b[6]))]
            bas.append([i, int(b[1]), int(b[2]), int(b[3]), 0, pe, pc, 0])
    norb = len(bas)
    aux = {}
    for l in range(aux_lmax + 1):
        for k in range(6 - l):
            a = 0.25 * 2.2 ** k
            pe = len(env)
            env += [a, gto_norm(l, a)]
            aux[(l, k)] = (pe, pe + 1)
    for i in range(natm):
        for (l, k), (pe, pc) in aux.items():
            bas.append([i, l, 1, 1, 0, pe, pc, 0])
    return atm, np.array(bas, np.int32), np.array(env), norb
