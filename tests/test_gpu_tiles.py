"""GPU parity of the whole-job / tile driver (cintb200_int2e_sph_all_unique): every block of every tile is
compared element-wise with the oracle on small molecules, for the register kernels and for the generic
kernel in tile mode, single rank and 2-/3-rank column sharding."""
import numpy as np
import pytest
import oracle_util as ou
import libcint_b200 as cb

pytestmark = pytest.mark.gpu


def pinned_sinks(n, nbytes):
    """n pinned host buffers of nbytes (torch is only the allocator here)."""
    import torch
    return [torch.empty(max(1, nbytes // 8), dtype=torch.float64, pin_memory=True) for _ in range(n)]


def check_job(name, nranks=1, force_generic=False, chunk_bytes=1 << 28, max_quartets=None, tol=1e-12, omega=None, nsinks=2, cart=False):
    """Whole job through the host-tile path (ring of pinned sinks + callback): EVERY chunk is delivered to the host and its
    blocks are compared element-wise with the oracle; entries outside the loop (k > i) must arrive as zeros."""
    which, _ = ou.best()
    atm, bas, env = cb.load_fixture(name)
    if omega is not None:
        env = env.copy()
        env[8] = omega                      # PTR_RANGE_OMEGA: > 0 long range (erf), < 0 short range (erfc)
        if omega < 0:
            which = "port"                  # the port evaluates erfc as full - erf in extended precision roots
    nbas = len(bas)
    dims = [((int(b[1]) + 1) * (int(b[1]) + 2) // 2 if cart else 2 * int(b[1]) + 1) * int(b[3]) for b in bas]
    intor = "int2e_cart" if cart else "int2e_sph"
    total_q = 0
    worst = [0.0]
    rng = np.random.default_rng(1)
    sinks = pinned_sinks(nsinks, chunk_bytes)
    for rank in range(nranks):
        ctx = cb.Context(atm, bas, env)
        if force_generic:
            ctx.force_generic(True)
        seen = []

        def on_tile(g, tile, ctx=ctx, rank=rank):
            seen.append(g["chunk"])
            (ri, rj, rpos), (ck, cl, cpos) = ctx.job_maps(g["chunk"])
            rowof = {(int(ri[r]), int(rj[r])): r for r in np.nonzero(rpos == 0)[0]}
            colof = {(int(ck[q]), int(cl[q])): q for q in np.nonzero(cpos == 0)[0]}
            assert tile.shape == (g["nrows"], g["ncols"])
            pairs_i = [(i, j) for i in range(g["i0"], g["i1"]) for j in range(i + 1)]
            assert sorted(rowof) == pairs_i
            kets = sorted(colof)
            assert all(k_ < g["i1"] for (k_, l) in kets)
            if max_quartets and len(pairs_i) * len(kets) > max_quartets * 7 // 10:
                sel = rng.choice(len(pairs_i), max(1, max_quartets * 7 // 10 // max(1, len(kets))), replace=False)
                pairs_i = [pairs_i[s] for s in sel]
            for (i, j) in pairs_i:
                r = rowof[(i, j)]
                nb = dims[i] * dims[j]
                for (k_, l) in kets:
                    c = colof[(k_, l)]
                    nk = dims[k_] * dims[l]
                    got = tile[r:r + nb, c:c + nk]
                    if k_ > i:              # outside the loop of examples/time_c60.c:206: never evaluated, delivered as zeros
                        assert not got.any(), (name, rank, (i, j, k_, l))
                        continue
                    want, _ = ou.eval_tuple(which, intor, (i, j, k_, l), atm, bas, env)
                    err = np.abs(got - want.reshape((nb, nk), order="F")).max()
                    scale = max(1.0, np.abs(want).max())
                    assert err <= tol * scale, (name, rank, (i, j, k_, l), err)
                    worst[0] = max(worst[0], err / scale)

        st = ctx.all_unique_tiles([t.data_ptr() for t in sinks], on_tile, rank=rank, nranks=nranks, chunk_bytes=chunk_bytes, cart=cart)
        total_q += st[0]
        assert seen == sorted(set(seen)) and len(seen) <= int(st[9]), seen      # chunks arrive once each, in order
        if nranks == 1:
            assert seen == list(range(int(st[9]))), seen
        ctx.close()
    nq = sum((i + 1) * (i + 1) * (i + 2) // 2 for i in range(nbas))
    assert total_q == nq, (total_q, nq)
    return worst[0]


def test_tiles_c2h6_631g_register_kernels():
    check_job("c2h6_631g")


def test_tiles_c2h6_ccpvdz_register_kernels():
    check_job("c2h6_ccpvdz")


def test_tiles_c2h6_ccpvdz_generic_tile_mode():
    check_job("c2h6_ccpvdz", force_generic=True, max_quartets=20000)


def test_tiles_two_and_three_ranks():
    check_job("c2h6_ccpvdz", nranks=2, max_quartets=30000)
    check_job("c2h6_631g", nranks=3)


def test_tiles_c2h6_ccpvtz_mixed_kernels():
    # f shells: the s/p/d classes use the specialised kernels, everything with an f shell the generic kernel
    check_job("c2h6_ccpvtz", max_quartets=6000)


def test_tiles_range_separated_specialised_kernels():
    # config 5 through the whole-job driver: the RS instantiations of the register / cooperative kernels (long range:
    # one attenuated rule; short range: full + negated attenuated rule), every block against the oracle
    check_job("c2h6_ccpvdz", omega=0.3, max_quartets=40000)
    check_job("c2h6_ccpvdz", omega=-0.3, max_quartets=40000, tol=1e-11)
    check_job("c2h6_ccpvtz", omega=0.3, max_quartets=5000)


def test_range_separated_density_fitting_and_blocks():
    # long-range (erf) 3-centre integrals through the whole-job driver and the dense block calls (RS kernels incl. f / g)
    worst, kinds = check_df_job(2, omega=0.3, max_triples=2500)
    assert kinds <= {1, 2}, kinds
    rng = np.random.default_rng(23)
    atm, bas, env = cb.load_fixture("c2h6_ccpvdz")
    env = env.copy()
    env[8] = 0.4
    ctx = cb.Context(atm, bas, env)
    _check_block(ctx, atm, bas, env, (0, 14, 3, 20, 5, 28, 0, 9), "int2e_sph", 2500, rng)
    from libcint_b200.basis import c60_df_basis
    atm, bas, env, norb = c60_df_basis(max_atoms=2)
    env = env.copy()
    env[8] = 0.4
    ctx = cb.Context(atm, bas, env)
    _check_block(ctx, atm, bas, env, (norb, len(bas), norb, len(bas)), "int2c2e_sph", 1500, rng, tol=1e-11)


def test_tiles_cartesian_output_specialised_kernels():
    # int2e_cart (north star: part of the hot path) from the register / cooperative kernels with the c2s stages compiled out:
    # every block of every chunk against the oracle's int2e_cart, d shells (cc-pVDZ) and f shells (cc-pVTZ), 1 and 2 ranks
    check_job("c2h6_ccpvdz", cart=True, max_quartets=40000)
    check_job("c2h6_ccpvdz", cart=True, nranks=2, chunk_bytes=4_000_000, max_quartets=30000)
    check_job("c2h6_ccpvtz", cart=True, max_quartets=6000)
    # the launches really are specialised kernels
    atm, bas, env = cb.load_fixture("c2h6_ccpvdz")
    ctx = cb.Context(atm, bas, env)
    ctx.all_unique(cart=True)
    assert set(ctx.launch_rows()[:, 7].astype(int)) <= {1, 2}
    # dense Cartesian blocks and list-mode batches (kind = CART) agree with the oracle too
    rng = np.random.default_rng(5)
    _check_block(ctx, atm, bas, env, (0, 14, 3, 20, 5, 28, 0, 9), "int2e_cart", 1500, rng)
    q = rng.integers(0, len(bas), size=(4000, 4)).astype(np.int32)
    v, o, s, nz = ctx.int2e_batch(q, kind=cb.CART)
    which, _ = ou.best()
    for n in rng.choice(len(q), 300, replace=False):
        want, _ = ou.eval_tuple(which, "int2e_cart", q[n], atm, bas, env)
        assert np.abs(v[o[n]:o[n] + s[n]] - want).max() <= 1e-12 * max(1.0, np.abs(want).max()), q[n]


def test_tiles_multi_chunk():
    # tiny chunk budget -> many chunks through the ring of sinks; every chunk is verified by value (1, 2 and 3 sinks)
    check_job("c2h6_ccpvdz", chunk_bytes=3_000_000)          # the largest tile (one bra shell x all kets) is 2.45 MB
    check_job("c2h6_631g", chunk_bytes=400_000, nsinks=1)
    check_job("c2h6_631g", chunk_bytes=400_000, nsinks=3, nranks=2)


def test_c60_job_statistics_and_sample():
    atm, bas, env = cb.load_fixture("c60_ccpvdz")
    ctx = cb.Context(atm, bas, env)
    st = ctx.all_unique(chunk_bytes=8 << 30)
    assert st[0] == 1023783775                       # shell quartets of examples/time_c60.c:200-207
    assert abs(st[1] - 6.3086e10) / 6.3086e10 < 1e-4  # integrals actually produced
    which, _ = ou.best()
    nch = int(st[9])
    tile, g = ctx.chunk(nch - 1)
    dims = [(2 * int(b[1]) + 1) * int(b[3]) for b in bas]
    rng = np.random.default_rng(8)
    for _ in range(300):
        i = int(rng.integers(g["i0"], g["i1"]))
        j = int(rng.integers(0, i + 1))
        k = int(rng.integers(0, i + 1))
        l = int(rng.integers(0, k + 1))
        r, _ = ctx.pair_offsets(i, j)
        _, c = ctx.pair_offsets(k, l)
        want, _ = ou.eval_tuple(which, "int2e_sph", (i, j, k, l), atm, bas, env)
        nb, nk = dims[i] * dims[j], dims[k] * dims[l]
        got = tile[r - g["row0"]:r - g["row0"] + nb, c:c + nk]
        assert np.abs(got - want.reshape((nb, nk), order="F")).max() <= 1e-12 * max(1.0, np.abs(want).max()), (i, j, k, l)


def _dimer(name, shift):
    """Two copies of a fixture molecule, the second shifted by `shift` Bohr along x (far apart -> most inter-molecular
    quartets are negligible)."""
    atm, bas, env = cb.load_fixture(name)
    natm, nbas = len(atm), len(bas)
    env2 = list(env)
    atm2 = np.vstack([atm, atm]).astype(np.int32)
    for i in range(natm):
        atm2[natm + i, 1] = len(env2)
        x, y, z = env[atm[i, 1]:atm[i, 1] + 3]
        env2 += [x + shift, y, z]
    bas2 = np.vstack([bas, bas]).astype(np.int32)
    bas2[nbas:, 0] += natm
    return atm2, bas2, np.array(env2)


def test_schwarz_bounds_and_screened_job():
    which, _ = ou.best()
    atm, bas, env = _dimer("c2h6_631g", 40.0)
    nbas = len(bas)
    dims = [(2 * int(b[1]) + 1) * int(b[3]) for b in bas]
    ctx = cb.Context(atm, bas, env)
    q = ctx.schwarz_bounds()
    assert len(q) == nbas * (nbas + 1) // 2
    rng = np.random.default_rng(2)
    # the bound holds: max|(ij|kl)| <= q_ij q_kl
    quart = rng.integers(0, nbas, (300, 4)).astype(np.int32)
    v, o, s, _ = ctx.int2e_batch(quart)
    for n, (i, j, k, l) in enumerate(quart):
        pij = max(i, j) * (max(i, j) + 1) // 2 + min(i, j)
        pkl = max(k, l) * (max(k, l) + 1) // 2 + min(k, l)
        assert np.abs(v[o[n]:o[n] + s[n]]).max() <= q[pij] * q[pkl] * (1 + 1e-10) + 1e-300
    # whole job with screening (default 1e-15): identical to the oracle within the parity tolerance, and genuinely
    # negligible blocks come out as exact zeros
    ctx.all_unique(chunk_bytes=1 << 30)
    tile, g = ctx.chunk(0)
    zeros = 0
    for _ in range(1500):
        i = int(rng.integers(0, nbas)); j = int(rng.integers(0, i + 1))
        k = int(rng.integers(0, i + 1)); l = int(rng.integers(0, k + 1))
        r, _ = ctx.pair_offsets(i, j)
        _, c = ctx.pair_offsets(k, l)
        want, _ = ou.eval_tuple(which, "int2e_sph", (i, j, k, l), atm, bas, env)
        nb, nk = dims[i] * dims[j], dims[k] * dims[l]
        got = tile[r:r + nb, c:c + nk]
        assert np.abs(got - want.reshape((nb, nk), order="F")).max() <= 1e-12 * max(1.0, np.abs(want).max()), (i, j, k, l)
        zeros += int(np.all(got == 0))
    assert zeros > 50                      # inter-molecular charge clouds do not overlap at 40 Bohr
    # switching the screening off changes nothing beyond 1e-15
    ctx2 = cb.Context(atm, bas, env)
    ctx2.set_schwarz_threshold(0.0)
    ctx2.all_unique(chunk_bytes=1 << 30)
    tile2, _ = ctx2.chunk(0)
    # compare the entries inside the loop (k <= i): the others are never written and hold whatever the buffer held before
    (ri, _, _), (ck, _, _) = ctx2.job_maps(0)
    valid = ck[None, :] <= ri[:, None]
    assert np.abs(np.where(valid, tile - tile2, 0.0)).max() < 1e-14


def test_c60_full_size_bench_configuration():
    # BASELINE.json configs[1] exactly as bench.py runs it (80 GB tile buffer, 7 chunks): sampled blocks of the last
    # chunk against the oracle, plus the size-independent property (ij|kl) == (kl|ij)^T between different blocks
    which, _ = ou.best()
    atm, bas, env = cb.load_fixture("c60_ccpvdz")
    ctx = cb.Context(atm, bas, env)
    st = ctx.all_unique(chunk_bytes=80 << 30)
    assert st[0] == 1023783775
    geom = np.zeros(8, dtype=np.int64)
    ctx.lib.cintb200_debug_chunk(ctx.handle, int(st[9]) - 1, None, 0, geom.ctypes.data_as(__import__("ctypes").c_void_p))
    i0, i1 = int(geom[0]), int(geom[1])
    dims = [(2 * int(b[1]) + 1) * int(b[3]) for b in bas]
    rng = np.random.default_rng(80)
    for n in range(400):
        i = int(rng.integers(i0, i1)); j = int(rng.integers(0, i + 1))
        k = int(rng.integers(0, i + 1)); l = int(rng.integers(0, k + 1))
        if n % 4 == 0:                       # force diagonal kets (k inside the chunk's own shell range)
            k = int(rng.integers(i0, i + 1)); l = int(rng.integers(0, k + 1))
        r, _ = ctx.pair_offsets(i, j)
        _, c = ctx.pair_offsets(k, l)
        nb, nk = dims[i] * dims[j], dims[k] * dims[l]
        got = ctx.block(r, c, nb, nk)
        want, _ = ou.eval_tuple(which, "int2e_sph", (i, j, k, l), atm, bas, env)
        assert np.abs(got - want.reshape((nb, nk), order="F")).max() <= 1e-12 * max(1.0, np.abs(want).max()), (i, j, k, l)
        if k >= i0 and i <= k:               # i == k: the transposed block (kl|ij) with i <= k is also part of the job
            r2, _ = ctx.pair_offsets(k, l)
            _, c2 = ctx.pair_offsets(i, j)
            tr = ctx.block(r2, c2, nk, nb)
            assert np.abs(tr - got.T).max() < 1e-13


def check_df_job(max_atoms, nranks=1, force_generic=False, chunk_bytes=1 << 30, max_triples=None, aux_lmax=4, omega=None):
    """int3c2e whole-job driver (cintb200_int3c2e_sph_all) on the first atoms of the config-3 stand-in: every block of
    the last tile against the oracle, triple and integral counts over all chunks and ranks."""
    from libcint_b200.basis import c60_df_basis
    which, _ = ou.best()
    atm, bas, env, norb = c60_df_basis(max_atoms=max_atoms, aux_lmax=aux_lmax)
    if omega is not None:
        env = env.copy()
        env[8] = omega
    nbas = len(bas)
    dims = [(2 * int(b[1]) + 1) * int(b[3]) for b in bas]
    rng = np.random.default_rng(5)
    tot_triples, tot_ints, worst, kinds = 0, 0, 0.0, set()
    for rank in range(nranks):
        ctx = cb.Context(atm, bas, env)
        if force_generic:
            ctx.force_generic(True)
        st = ctx.int3c2e_all(norb, rank=rank, nranks=nranks, chunk_bytes=chunk_bytes)
        tot_triples += st[0]
        tot_ints += st[1]
        kinds |= set(int(r[7]) for r in ctx.launch_rows())
        nch = int(st[9])
        tile, g = ctx.chunk(nch - 1)
        todo = [(i, j, k) for i in range(g["i0"], g["i1"]) for j in range(i + 1) for k in range(norb, nbas)]
        if max_triples and len(todo) > max_triples:
            todo = [todo[s] for s in rng.choice(len(todo), max_triples, replace=False)]
        for (i, j, k) in todo:
            c = ctx.aux_offset(k)
            if c < 0:
                continue
            r, _ = ctx.pair_offsets(i, j)
            r -= g["row0"]
            want, _ = ou.eval_tuple(which, "int3c2e_sph", (i, j, k), atm, bas, env)
            nb = dims[i] * dims[j]
            got = tile[r:r + nb, c:c + dims[k]]
            err = np.abs(got - want.reshape((nb, dims[k]), order="F")).max()
            scale = max(1.0, np.abs(want).max())
            assert err <= (1e-12 if omega is None else 1e-11) * scale, (rank, (i, j, k), err, scale)
            worst = max(worst, err / scale)
        ctx.close()
    assert tot_triples == norb * (norb + 1) // 2 * (nbas - norb)
    nao, naux = sum(dims[:norb]), sum(dims[norb:])
    npairs_ao = sum(dims[i] * dims[j] for i in range(norb) for j in range(i + 1))
    assert tot_ints == npairs_ao * naux, (tot_ints, nao, naux)
    return worst, kinds


def test_df_tiles_specialised_kernels():
    # s..f orbital shells x s..g auxiliary shells: register and cooperative kernels only (no generic launch)
    worst, kinds = check_df_job(2)
    assert kinds <= {1, 2}, kinds


def test_df_tiles_generic_and_sharded():
    check_df_job(2, force_generic=True, max_triples=1500)
    check_df_job(2, nranks=2, max_triples=3000)
    check_df_job(3, nranks=3, chunk_bytes=300_000, max_triples=2000)


def test_df_full_size_bench_configuration():
    # BASELINE.json configs[2] at full size (stand-in auxiliary basis, see libcint_b200/basis.py:c60_df_basis): the whole
    # density-fitting job exactly as bench.py runs it (one 62 GB tile), sampled blocks against the oracle, the symmetry
    # (ij|k) == (ji|k)^T between the diagonal-pair blocks, and the 2-GPU column shard of the same job
    from libcint_b200.basis import c60_df_basis
    which, _ = ou.best()
    atm, bas, env, norb = c60_df_basis()
    nbas = len(bas)
    ctx = cb.Context(atm, bas, env)
    st = ctx.int3c2e_all(norb, chunk_bytes=80 << 30)
    assert st[0] == 175284000 and st[1] == 7795008000 and int(st[9]) == 1
    assert set(int(r[7]) for r in ctx.launch_rows()) <= {1, 2}          # specialised kernels only
    dims = [(2 * int(b[1]) + 1) * int(b[3]) for b in bas]
    rng = np.random.default_rng(81)
    tol = 1e-11 if which == "ref" else 1e-12          # the reference itself carries ~1e-12 relative noise in the h/g Rys roots
    for n in range(300):
        i = int(rng.integers(0, norb)); j = int(rng.integers(0, i + 1)); k = int(rng.integers(norb, nbas))
        r, _ = ctx.pair_offsets(i, j)
        c = ctx.aux_offset(k)
        nb = dims[i] * dims[j]
        got = ctx.block(r, c, nb, dims[k])
        want, _ = ou.eval_tuple(which, "int3c2e_sph", (i, j, k), atm, bas, env)
        assert np.abs(got - want.reshape((nb, dims[k]), order="F")).max() <= tol * max(1.0, np.abs(want).max()), (i, j, k)
        if i == j:                                        # (ii|k) block is symmetric in its two orbital indices
            blk = got.reshape((dims[i], dims[i], dims[k]), order="F")
            assert np.abs(blk - blk.transpose(1, 0, 2)).max() < 1e-13
    ctx.close()
    ctx = cb.Context(atm, bas, env)
    st2 = ctx.int3c2e_all(norb, rank=1, nranks=2, chunk_bytes=80 << 30)
    assert abs(st2[1] * 2 - st[1]) <= 0.02 * st[1]
    for n in range(60):
        i = int(rng.integers(0, norb)); j = int(rng.integers(0, i + 1)); k = int(rng.integers(norb, nbas))
        c = ctx.aux_offset(k)
        if c < 0:
            continue
        r, _ = ctx.pair_offsets(i, j)
        got = ctx.block(r, c, dims[i] * dims[j], dims[k])
        want, _ = ou.eval_tuple(which, "int3c2e_sph", (i, j, k), atm, bas, env)
        assert np.abs(got - want.reshape(got.shape, order="F")).max() <= tol * max(1.0, np.abs(want).max()), (i, j, k)


def _check_block(ctx, atm, bas, env, sl, name, nsample, rng, tol=1e-12):
    """dense shell-slice tensor against the oracle, quartet by quartet (all of them, or a random sample)"""
    which, _ = ou.best()
    nc = len(sl) // 2
    cart = name.endswith("cart")
    arr, st = (ctx.int2e_block(sl, cart=cart) if nc == 4 else ctx.int3c2e_block(sl, cart=cart) if nc == 3 else ctx.int2c2e_block(sl, cart=cart))
    ao = np.concatenate([[0], np.cumsum([((int(b[1]) + 1) * (int(b[1]) + 2) // 2 if cart else 2 * int(b[1]) + 1) * int(b[3]) for b in bas])])
    ranges = [range(sl[2 * m], sl[2 * m + 1]) for m in range(nc)]
    import itertools
    tuples = list(itertools.product(*ranges))
    assert st[0] == len(tuples) and st[1] == arr.size
    if nsample and len(tuples) > nsample:
        tuples = [tuples[n] for n in rng.choice(len(tuples), nsample, replace=False)]
    for sh in tuples:
        want, _ = ou.eval_tuple(which, name, sh, atm, bas, env)
        idx = tuple(slice(int(ao[s] - ao[sl[2 * m]]), int(ao[s + 1] - ao[sl[2 * m]])) for m, s in enumerate(sh))
        got = arr[idx]
        want = want.reshape(got.shape, order="F")
        assert np.abs(got - want).max() <= tol * max(1.0, np.abs(want).max()), (sl, sh)
    return st


def test_dense_shell_slice_blocks():
    # cintb200_int2e_sph_block / cintb200_int3c2e_sph_block: the dense sub-tensor a fill driver asks for, on the tile kernels
    rng = np.random.default_rng(17)
    atm, bas, env = cb.load_fixture("c2h6_ccpvdz")
    nb = len(bas)
    ctx = cb.Context(atm, bas, env)
    st = _check_block(ctx, atm, bas, env, (0, nb, 0, nb, 0, nb, 0, nb), "int2e_sph", 4000, rng)      # the full 58^4 tensor
    assert set(int(r[7]) for r in ctx.launch_rows()) <= {1, 2}
    _check_block(ctx, atm, bas, env, (3, 11, 0, 5, 7, nb, 2, 9), "int2e_sph", 3000, rng)             # offset, non-square slices
    _check_block(ctx, atm, bas, env, (5, 6, 5, 6, 5, 6, 5, 6), "int2e_sph", 0, rng)                  # one quartet
    atm, bas, env = cb.load_fixture("c2h6_ccpvtz")                                                   # f shells: generic fallback mixed in
    ctx = cb.Context(atm, bas, env)
    _check_block(ctx, atm, bas, env, (0, 20, 10, 30, 5, 25, 30, 54), "int2e_sph", 2500, rng)
    from libcint_b200.basis import c60_df_basis
    atm, bas, env, norb = c60_df_basis(max_atoms=4)
    ctx = cb.Context(atm, bas, env)
    _check_block(ctx, atm, bas, env, (0, norb, 0, norb, norb, len(bas)), "int3c2e_sph", 4000, rng, tol=1e-11)
    _check_block(ctx, atm, bas, env, (9, 27, 0, 18, norb + 20, norb + 60), "int3c2e_sph", 3000, rng, tol=1e-11)
    m, _ = ctx.int2c2e_block((norb, len(bas), norb, len(bas)))                                      # the metric (P|Q)
    assert np.abs(m - m.T).max() < 1e-13 and np.linalg.eigvalsh(m).min() > 0
    _check_block(ctx, atm, bas, env, (norb, len(bas), norb + 7, len(bas) - 3), "int2c2e_sph", 2500, rng, tol=1e-11)
    # a larger block on the benchmark molecule: 20 x 20 x 150 x 150 shells = 9e6 quartets, 4.4 GB, in one call
    atm, bas, env = cb.load_fixture("c60_ccpvdz")
    ctx = cb.Context(atm, bas, env)
    _check_block(ctx, atm, bas, env, (100, 120, 20, 40, 0, 150, 150, 300), "int2e_sph", 1500, rng)
