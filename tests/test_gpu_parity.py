"""GPU parity tests (run with -m gpu on the B200 box): every call goes through the C ABI of
libcint_b200.so and is compared with the oracle (the compiled reference when oracle/_ref travelled with
the snapshot, else the plain-C port) and with committed golden vectors.

Tolerance (BASELINE.json north_star): 1e-12 absolute, 1e-10 relative for large values ->
|gpu - ref| <= 1e-12 * max(1, |ref|max of the block) is what we assert (tighter than 1e-10 relative)."""
import os
import numpy as np
import pytest
import oracle_util as ou
import libcint_b200 as cb
from libcint_b200.basis import reference_test_basis, class_sweep_basis, unique_quartets, c60_df_basis

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-12


def fp(v):
    return np.array([np.abs(v).sum(), (v * np.cos(np.arange(v.size))).sum()])


def split(vals, offs, sizes):
    return [vals[o:o + s] for o, s in zip(offs, sizes)]


def assert_blocks_close(got, want, shls, tol=TOL, what=""):
    worst = 0.0
    for g, w, s in zip(got, want, shls):
        err = np.abs(g - w).max() if g.size else 0.0
        scale = max(1.0, np.abs(w).max() if w.size else 0.0)
        assert err <= tol * scale, "%s shells %s: err %.3e (block max %.3e)" % (what, tuple(s), err, scale)
        worst = max(worst, err / scale)
    return worst


def test_known_answer_int2e_sph_through_dropin():
    # testsuite/test_cint.py:235-256,479 pointed at our cint2e_sph symbol, opt = NULL
    atm, bas, env = reference_test_basis()
    tot, cnt = 0.0, 0
    for l in range(8):
        for k in range(l + 1):
            for j in range(8):
                for i in range(j + 1):
                    v, rc = cb.int2e_sph((i, j, k, l), atm, bas, env)
                    tot += np.abs(v).sum()
                    cnt += v.size
    assert round(abs(tot - 56243.88080655417) / cnt ** .5, 8) == 0
    assert abs(tot - 56243.88080655417) < 1e-8


def test_known_answer_int3c2e_sph():
    # testsuite/test_3c2e.py:173-202,303
    atm, bas, env = reference_test_basis(with_fit_shells=True)
    ctx = cb.Context(atm, bas, env)
    tot, cnt = 0.0, 0
    for k in range(4):
        b2 = bas.copy()
        b2[8, 0] = bas[k, 0]
        c2 = cb.Context(atm, b2, env)
        t3 = [(i, j, k) for j in range(4) for i in range(4)]
        t4 = [(i, j, k, 8) for j in range(4) for i in range(4)]
        v3, o3, s3, _ = ctx.int3c2e_batch(t3)
        v4, o4, s4, _ = c2.int2e_batch(t4)
        for a, b in zip(split(v3, o3, s3), split(v4, o4, s4)):
            assert np.abs(a - b).max() <= 1e-12 * max(1.0, np.abs(b).max())
        tot += np.abs(v3).sum()
        cnt += v3.size
    assert round(abs(tot - 1586.350797347553) / cnt ** .5, 10) == 0


def test_known_answer_int2c2e_sph():
    # testsuite/test_3c2e.py:266-294,318 through our cint2c2e_sph / batch entry points
    which, _ = ou.best()
    atm, bas, env = reference_test_basis(with_fit_shells=True)
    ctx = cb.Context(atm, bas, env)
    t = np.array([(i, k) for k in range(8) for i in range(8)], np.int32)
    v, o, s, nz = ctx.int2c2e_batch(t)
    want = ou.eval_many(which, "int2c2e_sph", t, atm, bas, env)
    assert_blocks_close(split(v, o, s), want, t, what="int2c2e_sph")
    tot = sum(np.abs(b).sum() for b, sh in zip(split(v, o, s), t) if sh[0] < 4 and sh[1] < 4)
    assert abs(tot - 782.3104849606677) < 1e-9
    assert nz.all()
    v, o, s, nz = ctx.int2c2e_batch(t, kind=cb.CART)
    want = ou.eval_many(which, "int2c2e_cart", t, atm, bas, env)
    assert_blocks_close(split(v, o, s), want, t, what="int2c2e_cart")
    # drop-in symbol with dims embedding (out is a window of a larger matrix)
    big = np.full((40, 30), 7.0, order="F")
    d = cb.shell_dims(bas, (1, 3))
    sub, rc = cb.int2c2e_sph((1, 3), atm, bas, env, dims=(40, 30), out=big)
    ref_blk, _ = ou.eval_tuple(which, "int2c2e_sph", (1, 3), atm, bas, env)
    assert np.abs(big[:d[0], :d[1]] - ref_blk.reshape(d, order="F")).max() < 1e-12 and rc == 1
    assert (big[d[0]:, :] == 7.0).all() and (big[:, d[1]:] == 7.0).all()
    # range-separated metric: long-range + short-range == full
    for om in (0.3, -0.3):
        e2 = env.copy()
        e2[8] = om
        v2, o2, s2, _ = cb.Context(atm, bas, e2).int2c2e_batch(t)
        w2 = ou.eval_many(which, "int2c2e_sph", t, atm, bas, e2)
        assert_blocks_close(split(v2, o2, s2), w2, t, tol=1e-10 if om < 0 else TOL, what="int2c2e omega %g" % om)


def test_ip1_first_derivatives():
    # ( nabla i j | k l ) and ( nabla i j | k ): known answers of testsuite/test_cint.py:480 and testsuite/test_3c2e.py:304
    # through our batch entry points, element-wise parity with the oracle, Cartesian variant, drop-in symbol with dims
    which, _ = ou.best()
    atm, bas, env = reference_test_basis(with_fit_shells=True)
    ctx = cb.Context(atm, bas, env)
    q = np.array([(i, j, k, l) for l in range(8) for k in range(l + 1) for j in range(8) for i in range(j + 1)], np.int32)
    v, o, s, nz = ctx.int2e_ip1_batch(q)
    assert abs(np.abs(v).sum() - 115489.8647398112) < 1e-6
    rng = np.random.default_rng(12)
    sel = rng.choice(len(q), 250, replace=False)
    want = ou.eval_many(which, "int2e_ip1_sph", q[sel], atm, bas, env)
    got = split(v, o, s)
    assert_blocks_close([got[n] for n in sel], want, q[sel], what="int2e_ip1_sph")
    q2 = rng.integers(0, 8, (120, 4)).astype(np.int32)
    v2, o2, s2, _ = ctx.int2e_ip1_batch(q2, kind=cb.CART)
    assert_blocks_close(split(v2, o2, s2), ou.eval_many(which, "int2e_ip1_cart", q2, atm, bas, env), q2, what="int2e_ip1_cart")
    t3 = np.array([(i, j, k) for k in range(4) for j in range(4) for i in range(4)], np.int32)
    v3, o3, s3, _ = ctx.int3c2e_ip1_batch(t3)
    assert abs(np.abs(v3).sum() - 2242.052249221302) < 1e-8
    t3 = rng.integers(0, 8, (150, 3)).astype(np.int32)
    v3, o3, s3, _ = ctx.int3c2e_ip1_batch(t3)
    assert_blocks_close(split(v3, o3, s3), ou.eval_many(which, "int3c2e_ip1_sph", t3, atm, bas, env), t3, what="int3c2e_ip1_sph")
    # drop-in symbol, embedded with dims: component stride = product of dims (src/cart2sph.c:5340-5352)
    sh = (1, 2, 5, 3)
    d = cb.shell_dims(bas, sh)
    dims = (d[0] + 2, d[1] + 1, d[2], d[3] + 3)
    big = np.full(dims + (3,), 5.0, order="F")
    lib = cb.load_library()
    import ctypes
    cs, cd = (ctypes.c_int * 4)(*sh), (ctypes.c_int * 4)(*dims)
    a32, b32, e64 = np.ascontiguousarray(atm, np.int32), np.ascontiguousarray(bas, np.int32), np.ascontiguousarray(env)
    rc = lib.int2e_ip1_sph(big.ctypes.data_as(ctypes.c_void_p), cd, cs, a32.ctypes.data_as(ctypes.c_void_p), len(a32),
                           b32.ctypes.data_as(ctypes.c_void_p), len(b32), e64.ctypes.data_as(ctypes.c_void_p), None, None)
    ref_blk, _ = ou.eval_tuple(which, "int2e_ip1_sph", sh, atm, bas, env)
    ref_blk = ref_blk.reshape(tuple(d) + (3,), order="F")
    assert rc == 1 and np.abs(big[:d[0], :d[1], :d[2], :d[3], :] - ref_blk).max() < 1e-12 * max(1.0, np.abs(ref_blk).max())
    assert (big[d[0]:] == 5.0).all() and (big[:, d[1]:] == 5.0).all() and (big[:, :, :, d[3]:] == 5.0).all()


def test_df_gradient_integrals():
    # ( i j | nabla k ), ( nabla i | k ), ( i | nabla k ): testsuite/test_3c2e.py:305,319,320 + element-wise parity; the
    # density-fitting stand-in (f orbital shells, g auxiliary shells) through int3c2e_ip1 / int3c2e_ip2
    import itertools
    which, _ = ou.best()
    atm, bas, env = reference_test_basis(with_fit_shells=True)
    ctx = cb.Context(atm, bas, env)
    for fn, name, nc, ref in ((ctx.int3c2e_ip2_batch, "int3c2e_ip2_sph", 3, 1970.982483824248),
                              (ctx.int2c2e_ip1_batch, "int2c2e_ip1_sph", 2, 394.6515972715189),
                              (ctx.int2c2e_ip2_batch, "int2c2e_ip2_sph", 2, 394.6515972715189)):
        t = np.array(list(itertools.product(range(8), repeat=nc)), np.int32)
        v, o, s, _ = fn(t)
        blocks = split(v, o, s)
        tot = sum(np.abs(b).sum() for b, sh in zip(blocks, t) if (sh < 4).all())
        assert abs(tot - ref) < 1e-8, (name, tot)
        sel = np.random.default_rng(3).choice(len(t), min(len(t), 150), replace=False)
        assert_blocks_close([blocks[n] for n in sel], ou.eval_many(which, name, t[sel], atm, bas, env), t[sel], what=name)
    atm, bas, env, norb = c60_df_basis(max_atoms=6)
    naux = len(bas) - norb
    rng = np.random.default_rng(34)
    t = np.stack([rng.integers(0, norb, 200), rng.integers(0, norb, 200), norb + rng.integers(0, naux, 200)], 1).astype(np.int32)
    ctx = cb.Context(atm, bas, env)
    for fn, name in ((ctx.int3c2e_ip1_batch, "int3c2e_ip1_sph"), (ctx.int3c2e_ip2_batch, "int3c2e_ip2_sph")):
        v, o, s, _ = fn(t)
        assert_blocks_close(split(v, o, s), ou.eval_many(which, name, t, atm, bas, env), t, tol=1e-11 if which == "ref" else TOL, what="df " + name)


def test_golden_derivatives_and_metric():
    # committed golden vectors of the new integral types (tests/golden/derivs.npz) through the batch entry points
    g = np.load(os.path.join(GOLD, "derivs.npz"))
    atm, bas, env = reference_test_basis(with_fit_shells=True)
    ctx = cb.Context(atm, bas, env)
    calls = {"int2c2e_sph": (ctx.int2c2e_batch, cb.SPH), "int2c2e_ip1_sph": (ctx.int2c2e_ip1_batch, cb.SPH),
             "int2c2e_ip2_sph": (ctx.int2c2e_ip2_batch, cb.SPH), "int3c2e_ip1_sph": (ctx.int3c2e_ip1_batch, cb.SPH),
             "int3c2e_ip2_sph": (ctx.int3c2e_ip2_batch, cb.SPH), "int2e_ip1_sph": (ctx.int2e_ip1_batch, cb.SPH),
             "int2e_ip1_cart": (ctx.int2e_ip1_batch, cb.CART)}
    for name, (fn, kind) in calls.items():
        q, f = g["q_" + name], g["f_" + name]
        v, o, s, _ = fn(q, kind=kind)
        for n in range(len(q)):
            assert np.allclose(fp(v[o[n]:o[n] + s[n]]), f[n], rtol=1e-11, atol=1e-11), (name, tuple(q[n]))
    for n in range(int(g["nfull"])):
        name, sh, want = str(g["full%d_name" % n]), g["full%d_shls" % n], g["full%d_vals" % n]
        v, o, s, _ = calls[name][0](sh.reshape(1, -1), kind=calls[name][1])
        assert np.abs(v - want).max() <= 1e-12 * max(1.0, np.abs(want).max()), (name, tuple(sh))


def test_ip1_on_c60_sample():
    # gradient integrals on the benchmark molecule: contracted s shells, p/d shells, distant centres
    which, _ = ou.best()
    atm, bas, env = cb.load_fixture("c60_ccpvdz")
    ctx = cb.Context(atm, bas, env)
    rng = np.random.default_rng(99)
    q = rng.integers(0, 300, (300, 4)).astype(np.int32)
    v, o, s, _ = ctx.int2e_ip1_batch(q)
    assert_blocks_close(split(v, o, s), ou.eval_many(which, "int2e_ip1_sph", q, atm, bas, env), q, what="c60 int2e_ip1_sph")


def test_golden_testbasis_all_quartets():
    g = np.load(os.path.join(GOLD, "testbasis.npz"))
    atm, bas, env = reference_test_basis(with_fit_shells=True)
    ctx = cb.Context(atm, bas, env)
    v, o, s, nz = ctx.int2e_batch(g["q4"])
    f = np.array([fp(b) for b in split(v, o, s)])
    assert np.abs(f - g["f4"]).max() < 2e-10         # checksums sum up to 2e4 elements of size <= 300
    assert np.all(np.abs(f - g["f4"]) <= 1e-12 * np.maximum(1, np.abs(g["f4"])) * 50)
    v, o, s, nz = ctx.int3c2e_batch(g["q3"])
    f = np.array([fp(b) for b in split(v, o, s)])
    assert np.all(np.abs(f - g["f3"]) <= 1e-12 * np.maximum(1, np.abs(g["f3"])) * 50)
    v, o, s, nz = ctx.int2e_batch(g["qcart"], kind=cb.CART)
    f = np.array([fp(b) for b in split(v, o, s)])
    assert np.all(np.abs(f - g["fcart"]) <= 1e-12 * np.maximum(1, np.abs(g["fcart"])) * 50)
    env_lr = env.copy()
    env_lr[8] = float(g["omega_lr"])
    clr = cb.Context(atm, bas, env_lr)
    v, o, s, nz = clr.int2e_batch(g["qlr"])
    f = np.array([fp(b) for b in split(v, o, s)])
    assert np.all(np.abs(f - g["flr"]) <= 1e-12 * np.maximum(1, np.abs(g["flr"])) * 50)


def test_elementwise_vs_oracle_testbasis():
    which, _ = ou.best()
    atm, bas, env = reference_test_basis(with_fit_shells=True)
    ctx = cb.Context(atm, bas, env)
    rng = np.random.default_rng(11)
    q = rng.integers(0, 8, (600, 4)).astype(np.int32)
    v, o, s, nz = ctx.int2e_batch(q)
    want = ou.eval_many(which, "int2e_sph", q, atm, bas, env)
    assert_blocks_close(split(v, o, s), want, q, what="int2e_sph")
    v, o, s, nz = ctx.int2e_batch(q[:200], kind=cb.CART)
    want = ou.eval_many(which, "int2e_cart", q[:200], atm, bas, env)
    assert_blocks_close(split(v, o, s), want, q[:200], what="int2e_cart")
    t = rng.integers(0, 8, (300, 3)).astype(np.int32)      # zero-exponent fit shells are not valid 3c2e aux shells
    v, o, s, nz = ctx.int3c2e_batch(t)
    want = ou.eval_many(which, "int3c2e_sph", t, atm, bas, env)
    assert_blocks_close(split(v, o, s), want, t, what="int3c2e_sph")


def test_golden_c60_blocks():
    g = np.load(os.path.join(GOLD, "c60_blocks.npz"))
    atm, bas, env = cb.load_fixture("c60_ccpvdz")
    ctx = cb.Context(atm, bas, env)
    v, o, s, nz = ctx.int2e_batch(g["shls"])
    want = [g["values"][a:b] for a, b in zip(g["offsets"][:-1], g["offsets"][1:])]
    assert_blocks_close(split(v, o, s), want, g["shls"], what="c60")
    assert nz.all()


def test_c2h6_full_unique_tensor():
    # config 1 (examples/time_c2h6.c): every unique quartet of C2H6/6-31G and cc-pVDZ, element-wise
    which, _ = ou.best()
    for name in ("c2h6_631g", "c2h6_ccpvdz"):
        atm, bas, env = cb.load_fixture(name)
        q = unique_quartets(len(bas))
        if len(q) > 30000:
            q = q[np.random.default_rng(3).choice(len(q), 30000, replace=False)]
        ctx = cb.Context(atm, bas, env)
        v, o, s, nz = ctx.int2e_batch(q)
        want = ou.eval_many(which, "int2e_sph", q, atm, bas, env)
        assert_blocks_close(split(v, o, s), want, q, what=name)


def test_c2h6_large_bases_sampled():
    # config 1, larger bases of examples/time_c2h6.c: 6-311G**, cc-pVTZ (f), cc-pVQZ (g), sampled quartets
    which, _ = ou.best()
    for name, n in (("c2h6_6311gss", 4000), ("c2h6_ccpvtz", 4000), ("c2h6_ccpvqz", 3000)):
        atm, bas, env = cb.load_fixture(name)
        q = np.random.default_rng(4).integers(0, len(bas), (n, 4)).astype(np.int32)
        v, o, s, nz = cb.Context(atm, bas, env).int2e_batch(q)
        want = ou.eval_many(which, "int2e_sph", q, atm, bas, env)
        assert_blocks_close(split(v, o, s), want, q, what=name)


def test_class_sweep_s_to_g():
    # config 4: contracted (3 prim x 2 ctr) quartets over l = 0..4 on four centres
    which, _ = ou.best()
    atm, bas, env = class_sweep_basis(lmax=4)
    nsh = 5
    rng = np.random.default_rng(9)
    q = []
    for _ in range(160):
        ls = rng.integers(0, 5, 4)
        q.append([int(c * nsh + l) for c, l in zip(rng.permutation(4), ls)])
    q += [[4, 9, 14, 19], [19, 19, 19, 19], [3, 8, 13, 18]]       # (gg|gg) 4 centres, 1 centre, (ff|ff)
    q = np.array(q, np.int32)
    ctx = cb.Context(atm, bas, env)
    v, o, s, nz = ctx.int2e_batch(q)
    want = ou.eval_many(which, "int2e_sph", q, atm, bas, env)
    # the reference's own roots are only ~1e-11 accurate for nroots 7,9 ->
    # compare high classes against the extended-precision port with the north-star tolerance, and the
    # reference with 1e-10 relative
    tol = 1e-10 if which == "ref" else TOL
    assert_blocks_close(split(v, o, s), want, q, tol=tol, what="class sweep")
    wantp = ou.eval_many("port", "int2e_sph", q[-3:], atm, bas, env)
    assert_blocks_close(split(v, o, s)[-3:], wantp, q[-3:], tol=TOL, what="class sweep vs port")


def test_dims_embedding_and_return_value():
    atm, bas, env = reference_test_basis()
    s = (1, 5, 2, 0)
    d = ou.dims_of(bas, s)
    dims = [d[0] + 2, d[1] + 1, d[2] + 3, d[3] + 1]
    out = np.full(dims, 7.0, order="F")
    got, rc = cb.int2e_sph(s, atm, bas, env, dims=dims, out=out)
    ref, _ = ou.eval_tuple(ou.best()[0], "int2e_sph", s, atm, bas, env)
    ref = ref.reshape(d, order="F")
    assert rc == 1
    assert np.abs(got[:d[0], :d[1], :d[2], :d[3]] - ref).max() < 1e-12
    mask = np.ones(dims, bool)
    mask[:d[0], :d[1], :d[2], :d[3]] = False
    assert np.all(got[mask] == 7.0)              # only the addressed sub-block is written
    # far-apart centres: everything screened -> zero block, return 0 (src/cint2e.c:861-865)
    atm2, bas2, env2 = reference_test_basis()
    env2[atm2[3, 1]] += 200.0
    got, rc = cb.int2e_sph((3, 0, 3, 0), atm2, bas2, env2)
    assert rc == 0 and np.all(got == 0)


def test_symmetry_properties():
    # size-independent properties: (ij|kl) == (ji|kl)^T == (kl|ij)
    atm, bas, env = cb.load_fixture("c60_ccpvdz")
    ctx = cb.Context(atm, bas, env)
    rng = np.random.default_rng(21)
    q = rng.integers(0, 300, (200, 4)).astype(np.int32)
    v0, o0, s0, _ = ctx.int2e_batch(q)
    v1, o1, s1, _ = ctx.int2e_batch(q[:, [1, 0, 2, 3]])
    v2, o2, s2, _ = ctx.int2e_batch(q[:, [2, 3, 0, 1]])
    for n, s in enumerate(q):
        d = ou.dims_of(bas, s)
        a = v0[o0[n]:o0[n] + s0[n]].reshape(d, order="F")
        b = v1[o1[n]:o1[n] + s1[n]].reshape([d[1], d[0], d[2], d[3]], order="F")
        c = v2[o2[n]:o2[n] + s2[n]].reshape([d[2], d[3], d[0], d[1]], order="F")
        assert np.abs(a - b.transpose(1, 0, 2, 3)).max() < 1e-13
        assert np.abs(a - c.transpose(2, 3, 0, 1)).max() < 1e-13


def test_range_separated_coulomb():
    # config 5: env[PTR_RANGE_OMEGA] = +omega (erf, long range) / -omega (erfc, short range).
    # Properties from testsuite/test_cint.py:258-302 (LR + SR == full, checked there to 1e-9) plus element-wise
    # parity with the oracle.  Our SR is evaluated as full - LR with one 2N-point rule, so its absolute error
    # scales with the FULL integral: tolerance 1e-12 * max(1, |full| of the block).
    which, _ = ou.best()
    cases = [("testbasis", reference_test_basis()), ("c2h6_ccpvtz", cb.load_fixture("c2h6_ccpvtz"))]
    for name, (atm, bas, env) in cases:
        nb = len(bas)
        rng = np.random.default_rng(17)
        q = rng.integers(0, nb, (250, 4)).astype(np.int32)
        full, o, s, _ = cb.Context(atm, bas, env).int2e_batch(q)
        for omega in (0.3, 0.8):
            env_lr, env_sr = env.copy(), env.copy()
            env_lr[8], env_sr[8] = omega, -omega
            lr, _, _, _ = cb.Context(atm, bas, env_lr).int2e_batch(q)
            sr, _, _, _ = cb.Context(atm, bas, env_sr).int2e_batch(q)
            assert np.abs(lr + sr - full).max() < 1e-11, (name, omega)
            want_lr = ou.eval_many(which, "int2e_sph", q[:120], atm, bas, env_lr)
            want_sr = ou.eval_many("port", "int2e_sph", q[:120], atm, bas, env_sr)
            for n in range(120):
                blk = slice(o[n], o[n] + s[n])
                scale = max(1.0, np.abs(full[blk]).max())
                assert np.abs(lr[blk] - want_lr[n]).max() <= 1e-12 * scale, (name, omega, "lr", tuple(q[n]))
                assert np.abs(sr[blk] - want_sr[n]).max() <= 1e-12 * scale, (name, omega, "sr", tuple(q[n]))
    # whole-job driver with omega != 0: RS instantiations of the tile kernels
    atm, bas, env = cb.load_fixture("c2h6_631g")
    env = env.copy()
    env[8] = -0.5
    ctx = cb.Context(atm, bas, env)
    ctx.all_unique(chunk_bytes=1 << 30)
    tile, g = ctx.chunk(0)
    for (i, j, k, l) in [(5, 3, 4, 1), (21, 20, 7, 7), (10, 0, 10, 0)]:
        r, _ = ctx.pair_offsets(i, j)
        _, c = ctx.pair_offsets(k, l)
        want, _ = ou.eval_tuple("port", "int2e_sph", (i, j, k, l), atm, bas, env)
        d = ou.dims_of(bas, (i, j, k, l))
        got = tile[r:r + d[0] * d[1], c:c + d[2] * d[3]]
        assert np.abs(got - want.reshape(got.shape, order="F")).max() < 1e-12


def test_class_sweep_with_h_shells():
    # config 4 upper end: classes containing h shells (nroots up to 11), contracted 3 prim x 2 ctr
    atm, bas, env = class_sweep_basis(lmax=5)
    nsh = 6
    q = np.array([[5, 6 + 0, 12 + 5, 18 + 0],      # (hs|hs)  nroots 6
                  [5, 6 + 5, 12 + 0, 18 + 0],      # (hh|ss)  nroots 6
                  [5, 6 + 1, 12 + 2, 18 + 2],      # (hp|dd)  nroots 6
                  [5, 6 + 4, 12 + 3, 18 + 5],      # (hg|fh)  nroots 9
                  [4, 6 + 5, 12 + 5, 18 + 4]],     # (gh|hg)  nroots 10
                 np.int32)
    ctx = cb.Context(atm, bas, env)
    v, o, s, nz = ctx.int2e_batch(q)
    want = ou.eval_many("port", "int2e_sph", q, atm, bas, env)       # extended-precision roots: tight tolerance
    assert_blocks_close(split(v, o, s), want, q, tol=TOL, what="h classes vs port")
    if ou.ref() is not None:                                          # the reference's roots: ~1e-11 for nroots >= 7
        wantr = ou.eval_many("ref", "int2e_sph", q, atm, bas, env)
        assert_blocks_close(split(v, o, s), wantr, q, tol=1e-10, what="h classes vs reference")
    # one uncontracted (hh|hh) quartet: nroots 11, 14641 spherical integrals
    atm1, bas1, env1 = class_sweep_basis(lmax=5, nprim=1, nctr=1)
    q1 = np.array([[5, 11, 17, 23]], np.int32)
    v1, o1, s1, _ = cb.Context(atm1, bas1, env1).int2e_batch(q1)
    w1 = ou.eval_many("port", "int2e_sph", q1, atm1, bas1, env1)
    assert_blocks_close(split(v1, o1, s1), w1, q1, tol=TOL, what="(hh|hh)")


def test_int3c2e_density_fitting_sample():
    # config 3 stand-in: C60 geometry, carbon cc-pVTZ orbital shells, even-tempered s..g auxiliary shells
    which, _ = ou.best()
    atm, bas, env, norb = c60_df_basis(max_atoms=12)
    naux = len(bas) - norb
    rng = np.random.default_rng(33)
    t = np.stack([rng.integers(0, norb, 400), rng.integers(0, norb, 400), norb + rng.integers(0, naux, 400)], 1).astype(np.int32)
    ctx = cb.Context(atm, bas, env)
    v, o, s, nz = ctx.int3c2e_batch(t)
    want = ou.eval_many(which, "int3c2e_sph", t, atm, bas, env)
    assert_blocks_close(split(v, o, s), want, t, tol=1e-11 if which == "ref" else TOL, what="int3c2e df")
    v, o, s, nz = ctx.int3c2e_batch(t[:100], kind=cb.CART)
    want = ou.eval_many(which, "int3c2e_cart", t[:100], atm, bas, env)
    assert_blocks_close(split(v, o, s), want, t[:100], tol=1e-11 if which == "ref" else TOL, what="int3c2e cart")


def test_dropin_calls_from_concurrent_threads():
    # SURVEY 8b "Threading": the reference is called concurrently from OpenMP threads with one shared, read-only CINTOpt
    # (examples/time_c60.c:196-219).  Same pattern here: 8 host threads issue drop-in calls against one optimizer object
    # (ctypes releases the GIL during the call); every result must equal the oracle's.
    import ctypes
    import threading
    which, _ = ou.best()
    atm, bas, env = reference_test_basis()
    lib = cb.load_library()
    a32, b32, e64 = np.ascontiguousarray(atm, np.int32), np.ascontiguousarray(bas, np.int32), np.ascontiguousarray(env)
    pa, pb, pe = (x.ctypes.data_as(ctypes.c_void_p) for x in (a32, b32, e64))
    opt = ctypes.c_void_p()
    lib.cint2e_sph_optimizer(ctypes.byref(opt), pa, len(a32), pb, len(b32), pe)
    assert opt.value
    rng = np.random.default_rng(21)
    work = [[tuple(int(v) for v in rng.integers(0, 8, 4)) for _ in range(40)] for _ in range(8)]
    results = [None] * 8

    def run(n):
        out = []
        for sh in work[n]:
            d = cb.shell_dims(b32, sh)
            buf = np.zeros(int(np.prod(d)))
            cs = (ctypes.c_int * 4)(*sh)
            rc = lib.cint2e_sph(buf.ctypes.data_as(ctypes.c_void_p), cs, pa, len(a32), pb, len(b32), pe, opt)
            out.append((buf, rc))
        results[n] = out

    th = [threading.Thread(target=run, args=(n,)) for n in range(8)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for n in range(8):
        for sh, (buf, rc) in zip(work[n], results[n]):
            want, r0 = ou.eval_tuple(which, "int2e_sph", sh, atm, bas, env)
            assert rc == r0 and np.abs(buf - want).max() <= TOL * max(1.0, np.abs(want).max()), (n, sh)
    lib.CINTdel_optimizer(ctypes.byref(opt))
    assert not opt.value


def test_large_lists_device_side_bookkeeping():
    """Lists of >= 4096 tuples are keyed, sorted and turned into work items ON THE DEVICE (csrc/listdev.cu): random C60 quartets,
    packed and caller-given offsets, host and device output, Cartesian output, 3- and 2-centre tuples -- sampled blocks against
    the oracle and the nonzero flags against the reference's return value."""
    which, _ = ou.best()
    atm, bas, env = cb.load_fixture("c60_ccpvdz")
    ctx = cb.Context(atm, bas, env)
    rng = np.random.default_rng(11)
    n = 20000
    q = rng.integers(0, len(bas), size=(n, 4)).astype(np.int32)
    q[:500, 2:] = q[0, 2:]                      # a long run of one ket and ...
    q[500:900, :2] = q[1, :2]                   # ... one bra with many kets
    v, o, s, nz = ctx.int2e_batch(q)
    assert nz.min() >= 0 and nz.max() == 1
    pick = np.concatenate([np.arange(0, 40), rng.choice(n, 260, replace=False)])
    for t in pick:
        want, ret = ou.eval_tuple(which, "int2e_sph", q[t], atm, bas, env)
        assert np.abs(v[o[t]:o[t] + s[t]] - want).max() <= TOL * max(1.0, np.abs(want).max()), (t, q[t])
        assert nz[t] == ret or np.abs(want).max() < 1e-30
    # caller-given offsets (blocks in reverse order with gaps) into a device buffer
    import torch
    off2 = (np.cumsum((s + 3)[::-1])[::-1] - (s + 3)).astype(np.uint64)
    dbuf = torch.full((int((off2 + s.astype(np.uint64)).max()),), -7.0, dtype=torch.float64, device="cuda")
    ctx.int2e_batch(q, out_off=off2, device_ptr=dbuf.data_ptr())
    h = dbuf.cpu().numpy()
    for t in pick[:80]:
        assert np.array_equal(h[int(off2[t]):int(off2[t]) + s[t]], v[o[t]:o[t] + s[t]]), t
    assert h[int(off2[5]) + s[5]] == -7.0        # the gaps stay untouched
    # Cartesian output
    vc, oc, sc, _ = ctx.int2e_batch(q[:6000], kind=cb.CART)
    for t in pick[pick < 6000][:60]:
        want, _ = ou.eval_tuple(which, "int2e_cart", q[t], atm, bas, env)
        assert np.abs(vc[oc[t]:oc[t] + sc[t]] - want).max() <= TOL * max(1.0, np.abs(want).max()), (t, q[t])
    ctx.close()
    # 3- and 2-centre tuples of the density-fitting stand-in
    from libcint_b200.basis import c60_df_basis
    atm, bas, env, norb = c60_df_basis(max_atoms=6)
    ctx = cb.Context(atm, bas, env)
    q3 = np.column_stack([rng.integers(0, norb, 8000), rng.integers(0, norb, 8000), rng.integers(norb, len(bas), 8000)]).astype(np.int32)
    v3, o3, s3, _ = ctx.int3c2e_batch(q3)
    for t in rng.choice(8000, 150, replace=False):
        want, _ = ou.eval_tuple(which, "int3c2e_sph", q3[t], atm, bas, env)
        assert np.abs(v3[o3[t]:o3[t] + s3[t]] - want).max() <= 1e-11 * max(1.0, np.abs(want).max()), (t, q3[t])
    q2 = rng.integers(norb, len(bas), size=(5000, 2)).astype(np.int32)
    v2, o2, s2, _ = ctx.int2c2e_batch(q2)
    for t in rng.choice(5000, 150, replace=False):
        want, _ = ou.eval_tuple(which, "int2c2e_sph", q2[t], atm, bas, env)
        assert np.abs(v2[o2[t]:o2[t] + s2[t]] - want).max() <= 1e-11 * max(1.0, np.abs(want).max()), (t, q2[t])


def test_dense_block_generally_contracted_shells_virtual_pairs():
    """Dense shell-slice blocks over a basis in which EVERY shell carries two general contractions (class-sweep basis, s..f):
    pairs above s have no contracted kernel instantiation and enter the block as single-contraction virtual pairs
    (engine.cu:build_pairs, driver.cu:run_block) -- spherical and Cartesian output against the reference per tuple."""
    which, _ = ou.best()
    import itertools
    from libcint_b200.basis import class_sweep_basis
    atm, bas, env = class_sweep_basis(lmax=3, nctr=2)
    ctx = cb.Context(atm, bas, env)
    rng = np.random.default_rng(11)
    sl = (0, 8, 2, 8, 8, 12, 12, 16)
    for cart in (False, True):
        arr, st = ctx.int2e_block(sl, cart=cart)
        ao = np.concatenate([[0], np.cumsum([((int(b[1]) + 1) * (int(b[1]) + 2) // 2 if cart else 2 * int(b[1]) + 1) * int(b[3]) for b in bas])])
        tuples = list(itertools.product(*[range(sl[2 * m], sl[2 * m + 1]) for m in range(4)]))
        for sh in [tuples[n] for n in rng.choice(len(tuples), 250, replace=False)]:
            want, _ = ou.eval_tuple(which, "int2e_cart" if cart else "int2e_sph", sh, atm, bas, env)
            idx = tuple(slice(int(ao[s] - ao[sl[2 * m]]), int(ao[s + 1] - ao[sl[2 * m]])) for m, s in enumerate(sh))
            got = arr[idx]
            want = np.asarray(want).reshape(got.shape, order="F")
            assert np.abs(got - want).max() <= 1e-11 * max(1.0, np.abs(want).max()), (cart, sh, np.abs(got - want).max())
    ctx.close()


def test_ip1_dense_blocks_on_tile_kernels():
    """( nabla i j | k l ) and ( nabla i j | k ) over shell slices (cintb200_int2e_ip1_block): raised / lowered helper blocks on the
    specialised kernels + derivative and cart->sph on the dense tensor, against the reference's int2e_ip1 / int3c2e_ip1 per tuple."""
    which, _ = ou.best()
    rng = np.random.default_rng(3)
    import itertools
    for name, sl, kinds in (("c2h6_ccpvdz", (0, 14, 3, 20, 5, 28, 0, 9), (cb.SPH, cb.CART)), ("c2h6_ccpvtz", (20, 30, 0, 12, 30, 54, 5, 15), (cb.SPH,))):
        atm, bas, env = cb.load_fixture(name)
        ctx = cb.Context(atm, bas, env)
        for kind in kinds:
            cart = kind == cb.CART
            arr, st = ctx.ip1_block(sl, kind=kind)
            ao = np.concatenate([[0], np.cumsum([((int(b[1]) + 1) * (int(b[1]) + 2) // 2 if cart else 2 * int(b[1]) + 1) * int(b[3]) for b in bas])])
            tuples = list(itertools.product(*[range(sl[2 * m], sl[2 * m + 1]) for m in range(4)]))
            for sh in [tuples[n] for n in rng.choice(len(tuples), 500, replace=False)]:
                want, _ = ou.eval_tuple(which, "int2e_ip1_cart" if cart else "int2e_ip1_sph", sh, atm, bas, env)
                idx = tuple(slice(int(ao[s] - ao[sl[2 * m]]), int(ao[s + 1] - ao[sl[2 * m]])) for m, s in enumerate(sh))
                got = arr[idx]
                want = want.reshape(got.shape, order="F")
                assert np.abs(got - want).max() <= 1e-11 * max(1.0, np.abs(want).max()), (name, kind, sh, np.abs(got - want).max())
        ctx.close()
    from libcint_b200.basis import c60_df_basis
    atm, bas, env, norb = c60_df_basis(max_atoms=3)
    ctx = cb.Context(atm, bas, env)
    sl = (0, norb, 0, 12, norb, len(bas))
    arr, st = ctx.ip1_block(sl)
    ao = np.concatenate([[0], np.cumsum([(2 * int(b[1]) + 1) * int(b[3]) for b in bas])])
    tuples = list(itertools.product(*[range(sl[2 * m], sl[2 * m + 1]) for m in range(3)]))
    for sh in [tuples[n] for n in rng.choice(len(tuples), 400, replace=False)]:
        want, _ = ou.eval_tuple(which, "int3c2e_ip1_sph", sh, atm, bas, env)
        idx = tuple(slice(int(ao[s] - ao[sl[2 * m]]), int(ao[s + 1] - ao[sl[2 * m]])) for m, s in enumerate(sh))
        got = arr[idx]
        assert np.abs(got - want.reshape(got.shape, order="F")).max() <= 1e-11 * max(1.0, np.abs(want).max()), sh


def test_all_symmetry_unique_classes_s_to_h_segmented():
    """BASELINE config 4 with segmented shells (3 primitives, 1 contraction -- what real basis sets have above p): one quartet of
    EVERY symmetry-unique angular class (li>=lj, lk>=ll, ij>=kl; 231 classes, nroots 1..11) plus the transposed orientation of a
    sample, against the reference (1e-10: its own roots are ~1e-11 accurate for nroots >= 7) and the extended-precision port.
    Classes without a register / cooperative instantiation run on the wide kernel (csrc/kern_wide.cu) + the catch-all epilogue."""
    atm, bas, env = class_sweep_basis(lmax=5, nctr=1)
    nsh = 6
    pairs = [(a, b) for a in range(6) for b in range(a + 1)]
    classes = [(p[0], p[1], r[0], r[1]) for n, p in enumerate(pairs) for r in pairs[:n + 1]]
    assert len(classes) == 231
    q = [[0 * nsh + c[0], 1 * nsh + c[1], 2 * nsh + c[2], 3 * nsh + c[3]] for c in classes]
    q += [[2 * nsh + c[2], 3 * nsh + c[3], 0 * nsh + c[0], 1 * nsh + c[1]] for c in classes[::7]]       # (kl|ij): the other orientation
    q += [[1 * nsh + c[1], 0 * nsh + c[0], 3 * nsh + c[3], 2 * nsh + c[2]] for c in classes[3::11]]     # (ji|lk)
    q = np.array(q, np.int32)
    ctx = cb.Context(atm, bas, env)
    got = []
    for sh in q:                                    # 40 copies per call: list mode for the specialised classes, wide / catch-all else
        v, o, s, nz = ctx.int2e_batch(np.tile(sh, (40, 1)))
        assert np.array_equal(v[:s[0]], v[o[39]:o[39] + s[39]])
        got.append(v[:s[0]].copy())
    which, _ = ou.best()
    want = ou.eval_many(which, "int2e_sph", q, atm, bas, env)
    assert_blocks_close(got, want, q, tol=1e-10 if which == "ref" else TOL, what="all classes s..h")
    hi = [n for n, sh in enumerate(q) if sum(int(bas[s, 1]) for s in sh) >= 12][::3]
    wantp = ou.eval_many("port", "int2e_sph", q[hi], atm, bas, env)
    assert_blocks_close([got[n] for n in hi], wantp, q[hi], tol=TOL, what="high classes vs port")
    # long-range Coulomb and Cartesian output on high classes, 3-centre (ij|k) with g / h shells
    env2 = env.copy()
    env2[8] = 0.4
    ctx2 = cb.Context(atm, bas, env2)
    sel = np.array([q[n] for n in (230, 200, 150, 120)], np.int32)
    v, o, s, _ = ctx2.int2e_batch(sel)
    # (hh|hh): 11 roots, 63 001 Cartesian components folded through four h transforms -- rounding reaches 1.4e-12 on values ~ 1
    assert_blocks_close(split(v, o, s), ou.eval_many("port", "int2e_sph", sel, atm, bas, env2), sel, tol=3e-12, what="LR high classes")
    v, o, s, _ = ctx.int2e_batch(sel[1:], kind=cb.CART)
    assert_blocks_close(split(v, o, s), ou.eval_many(which, "int2e_cart", sel[1:], atm, bas, env), sel[1:], tol=1e-10, what="cart high classes")
    q3 = np.array([[5, 6 + 5, 12 + 4], [4, 6 + 4, 12 + 5], [5, 6 + 3, 12 + 5], [3, 6 + 3, 12 + 4]], np.int32)
    v, o, s, _ = ctx.int3c2e_batch(q3)
    assert_blocks_close(split(v, o, s), ou.eval_many(which, "int3c2e_sph", q3, atm, bas, env), q3, tol=1e-10, what="3-centre high classes")


def test_int3c2e_sph_ssc_dropin():
    """int3c2e_sph_ssc (src/cint3c2e.c:729): spherical orbital indices, Cartesian auxiliary index -- against the compiled reference."""
    if ou.ref() is None:
        pytest.skip("needs oracle/_ref (the port has no _ssc entry point)")
    from libcint_b200.basis import c60_df_basis
    atm, bas, env, norb = c60_df_basis(max_atoms=2)
    ctx = cb.Context(atm, bas, env)
    rng = np.random.default_rng(4)
    for _ in range(60):
        sh = (int(rng.integers(0, norb)), int(rng.integers(0, norb)), int(rng.integers(norb, len(bas))))
        got, rc = cb.int3c2e_sph_ssc(sh, atm, bas, env, opt=ctx)
        want, ret = ou.eval_tuple("ref", "int3c2e_sph_ssc", sh, atm, bas, env)
        assert rc == ret
        assert np.abs(got.ravel(order="F") - want).max() <= 1e-11 * max(1.0, np.abs(want).max()), sh
