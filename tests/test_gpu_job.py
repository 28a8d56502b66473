"""Whole-job value checks: EVERY block of a whole-job run is covered, through fingerprints that do not depend on chunking or
sharding (per-bra-pair sums S / A / F and the J / K matrices of a formula density), against goldens computed from the
UNMODIFIED reference by oracle/ref_golden.c (tools/make_golden_job.py -> tests/golden/job_*.npz).

Covered paths: device-resident tiles (checksum consumer), tiles delivered to the host through the ring of sinks (recomputed
with numpy from what the callback receives), the J/K digestion kernels, single rank and rank-sharded runs (partial results
summed like bench.py all-reduces them), at the bench's own chunk sizes for C60."""
import os
import numpy as np
import pytest
import oracle_util as ou
import libcint_b200 as cb
from test_gpu_tiles import pinned_sinks

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLD, "job_%s.npz" % name))


def assert_fingerprints(S, A, F, gold, what, rtol=2e-12):
    """|dS|, |dF| <= rtol * A (A = sum |v| of the pair: the scale rounding errors of a sum follow) and A itself to rtol."""
    scale = np.maximum(gold["A"], 1.0)
    for got, key in ((S, "S"), (A, "A"), (F, "F")):
        err = np.abs(got - gold[key]) / scale
        p = int(np.argmax(err))
        assert err[p] <= rtol, (what, key, "pair", p, float(err[p]), float(got[p]), float(gold[key][p]))


def sum_over_ranks(fn, nranks):
    tot = None
    for rank in range(nranks):
        parts = fn(rank)
        tot = parts if tot is None else [a + b for a, b in zip(tot, parts)]
    return tot


@pytest.mark.parametrize("name,chunk", [("c2h6_631g", 40_000), ("c2h6_ccpvdz", 300_000), ("c2h6_ccpvtz", 64 << 20)])
def test_checksums_every_chunk_small(name, chunk):
    atm, bas, env = cb.load_fixture(name)
    gold = golden(name)
    for nranks in (1, 3):
        def run(rank):
            ctx = cb.Context(atm, bas, env)
            ctx.set_checksums(True)
            st = ctx.all_unique(rank=rank, nranks=nranks, chunk_bytes=chunk)
            S, A, F = ctx.job_checksums()
            assert abs(st[3] - S.sum()) <= 1e-9 * max(1.0, A.sum())
            if name != "c2h6_ccpvtz":
                assert st[9] > 3              # really multi-chunk
            ctx.close()
            return [S, A, F]
        S, A, F = sum_over_ranks(run, nranks)
        assert_fingerprints(S, A, F, gold, (name, nranks))


def host_fingerprints(ctx, bas, sinks, **kw):
    """S and F recomputed on the HOST from the tiles the callback receives (A is skipped: needs a temporary per tile)."""
    dims = np.array([(2 * int(b[1]) + 1) * int(b[3]) for b in bas])
    ao = np.concatenate([[0], np.cumsum(dims)])
    npair = len(bas) * (len(bas) + 1) // 2
    h, g, _, _ = cb.job_weights(int(ao[-1]), int(dims.max()) ** 2)
    S, F = np.zeros(npair), np.zeros(npair)
    ntiles = [0]

    def on_tile(info, tile):
        ntiles[0] += 1
        (ri, rj, rpos), (ck, cl, cpos) = ctx.job_maps(info["chunk"])
        rp = ri.astype(np.int64) * (ri + 1) // 2 + rj
        gw = g[ao[ck] + cpos % dims[ck], ao[cl] + cpos // dims[ck]]
        nb = info["ncols_below"]
        # entries with K > I (only among the chunk's own kets, columns >= ncols_below) arrive as zeros, so one product of the
        # whole tile with [1, g] gives the in-loop row sums exactly
        if info["ncols"] > nb:
            assert not tile[:, nb:][ri[:, None] < ck[None, nb:]].any()
        w2 = np.asfortranarray(np.stack([np.ones(info["ncols"]), gw], axis=1))
        r2 = tile @ w2
        rs, rf = r2[:, 0], r2[:, 1]
        np.add.at(S, rp, rs)
        np.add.at(F, rp, h[rpos] * rf)

    st = ctx.all_unique_tiles([t.data_ptr() for t in sinks], on_tile, **kw)
    return S, F, st, ntiles[0]


def test_host_tiles_fingerprints_small():
    name = "c2h6_ccpvdz"
    atm, bas, env = cb.load_fixture(name)
    gold = golden(name)
    sinks = pinned_sinks(2, 3_000_000)
    for nranks in (1, 2):
        def run(rank):
            ctx = cb.Context(atm, bas, env)
            S, F, st, nt = host_fingerprints(ctx, bas, sinks, rank=rank, nranks=nranks, chunk_bytes=3_000_000)
            assert nt > 3 and st[5] > 0
            ctx.close()
            return [S, F]
        S, F = sum_over_ranks(run, nranks)
        scale = np.maximum(gold["A"], 1.0)
        assert (np.abs(S - gold["S"]) / scale).max() < 2e-12
        assert (np.abs(F - gold["F"]) / scale).max() < 2e-12


def assert_jk(vj, vk, gold, nao, what, rtol=1e-11):
    _, _, D, U = cb.job_weights(nao)
    assert np.abs(vj - vj.T).max() <= 1e-12 * np.abs(vj).max()
    assert np.abs(vk - vk.T).max() <= 1e-12 * np.abs(vk).max()
    for m, key in ((vj, "J"), (vk, "K")):
        sc = np.abs(gold[key + "U"]).max()
        assert np.abs(m @ U - gold[key + "U"]).max() <= rtol * sc, (what, key, np.abs(m @ U - gold[key + "U"]).max() / sc)
        assert np.abs(np.diag(m) - gold[key + "diag"]).max() <= rtol * np.abs(gold[key + "diag"]).max(), (what, key, "diag")
        assert abs(np.sum(m * D) - float(gold["tr%sD" % key])) <= rtol * abs(float(gold["tr%sD" % key])), (what, key, "trace")


def test_jk_dense_reference_631g():
    """J and K against the dense definition: the full (ab|cd) tensor from the oracle contracted with numpy."""
    name = "c2h6_631g"
    which, _ = ou.best()
    atm, bas, env = cb.load_fixture(name)
    nb = len(bas)
    dim = [(2 * int(b[1]) + 1) * int(b[3]) for b in bas]
    ao = np.concatenate([[0], np.cumsum(dim)])
    nao = int(ao[-1])
    T = np.zeros((nao,) * 4)
    for i in range(nb):
        for j in range(i + 1):
            for k in range(nb):
                for l in range(k + 1):
                    v, _ = ou.eval_tuple(which, "int2e_sph", (i, j, k, l), atm, bas, env)
                    blk = v.reshape((dim[i], dim[j], dim[k], dim[l]), order="F")
                    for (p, q, bb) in ((i, j, blk), (j, i, blk.transpose(1, 0, 2, 3))):
                        T[ao[p]:ao[p + 1], ao[q]:ao[q + 1], ao[k]:ao[k + 1], ao[l]:ao[l + 1]] = bb
                        T[ao[p]:ao[p + 1], ao[q]:ao[q + 1], ao[l]:ao[l + 1], ao[k]:ao[k + 1]] = bb.transpose(0, 1, 3, 2)
    _, _, D, _ = cb.job_weights(nao)
    J = np.einsum("abcd,cd->ab", T, D)
    K = np.einsum("abcd,bd->ac", T, D)
    for chunk in (0, 30_000):
        ctx = cb.Context(atm, bas, env)
        vj, vk, st = ctx.jk(D, chunk_bytes=chunk)
        assert np.abs(vj - J).max() <= 1e-12 * np.abs(J).max(), np.abs(vj - J).max()
        assert np.abs(vk - K).max() <= 1e-12 * np.abs(K).max(), np.abs(vk - K).max()
        vj2, _, _ = ctx.jk(D, chunk_bytes=chunk, with_k=False)
        assert np.abs(vj2 - J).max() <= 1e-12 * np.abs(J).max()
        ctx.close()


@pytest.mark.parametrize("name,chunk", [("c2h6_ccpvdz", 300_000), ("c2h6_ccpvtz", 64 << 20)])
def test_jk_vs_reference_golden(name, chunk):
    atm, bas, env = cb.load_fixture(name)
    gold = golden(name)
    nao = int(sum((2 * int(b[1]) + 1) * int(b[3]) for b in bas))
    _, _, D, _ = cb.job_weights(nao)
    for nranks in (1, 2):
        def run(rank):
            ctx = cb.Context(atm, bas, env)
            vj, vk, st = ctx.jk(D, rank=rank, nranks=nranks, chunk_bytes=chunk)
            ctx.close()
            return [vj, vk]
        vj, vk = sum_over_ranks(run, nranks)
        assert_jk(vj, vk, gold, nao, (name, nranks))


def test_c60_full_job_checksums_and_jk():
    """The benchmarked job itself (BASELINE config 2, all 1 023 783 775 quartets) at the bench's chunk size and at the
    end-to-end chunk size: every bra pair's fingerprint and the digested J / K against the reference goldens."""
    name = "c60_ccpvdz"
    atm, bas, env = cb.load_fixture(name)
    gold = golden(name)
    nao = 840
    _, _, D, _ = cb.job_weights(nao)
    ctx = cb.Context(atm, bas, env)
    ctx.set_checksums(True)
    for chunk_gb in (80, 16):
        st = ctx.all_unique(chunk_bytes=chunk_gb << 30)
        assert st[0] == 1023783775
        S, A, F = ctx.job_checksums()
        assert_fingerprints(S, A, F, gold, ("c60", chunk_gb))
        assert abs(st[3] - gold["S"].sum()) <= 1e-10 * gold["A"].sum()
    ctx.set_checksums(False)
    vj, vk, st = ctx.jk(D, chunk_bytes=80 << 30)
    assert_jk(vj, vk, gold, nao, "c60 jk")
    # 2-rank shard: partial fingerprints and partial J / K add up (what bench.py all-reduces)
    ctx.set_checksums(True)
    tot = None
    for rank in range(2):
        vj_r, vk_r, st = ctx.jk(D, rank=rank, nranks=2, chunk_bytes=80 << 30)
        parts = list(ctx.job_checksums()) + [vj_r, vk_r]
        tot = parts if tot is None else [a + b for a, b in zip(tot, parts)]
    assert_fingerprints(tot[0], tot[1], tot[2], gold, "c60 2 ranks")
    assert_jk(tot[3], tot[4], gold, nao, "c60 jk 2 ranks")
    ctx.close()


def test_c60_full_job_through_host_tiles():
    """The end-to-end path of bench.py (16 GB tiles through two pinned sinks): S and F of every bra pair recomputed on the host
    from the delivered tiles."""
    name = "c60_ccpvdz"
    atm, bas, env = cb.load_fixture(name)
    gold = golden(name)
    chunk = 12 << 30                         # one bra shell x all kets of C60 is an 11.9 GB tile
    sinks = pinned_sinks(2, chunk)
    ctx = cb.Context(atm, bas, env)
    S, F, st, nt = host_fingerprints(ctx, bas, sinks, chunk_bytes=chunk)
    assert nt == int(st[9]) and st[0] == 1023783775
    scale = np.maximum(gold["A"], 1.0)
    assert (np.abs(S - gold["S"]) / scale).max() < 2e-12
    assert (np.abs(F - gold["F"]) / scale).max() < 2e-12
    ctx.close()
