// List mode with the bookkeeping ON THE DEVICE: the batched entry points (cintb200_int2e_batch & co., include/cint_b200.h)
// take an arbitrary list of shell tuples.  driver.cu:list_mode_run turns such a list into work items for the specialised
// tile kernels on the host (OpenMP threads + a parallel sort: ~8e6 tuples/s, slower than the GPU evaluates them by two orders
// of magnitude).  Here the same steps run as kernels and cub primitives, so a call costs one upload of the shell list:
//   1. per tuple: pair ids, canonical orientation and strides, block size, class group, sort key      (list_key_kernel)
//   2. packed output offsets = exclusive scan of the block sizes                                        (cub::DeviceScan)
//   3. sort by (group | ket | ket orientation | primitive count of the bra)                            (cub::DeviceRadixSort)
//   4. runs of equal (group, ket, orientation) -> ket entries; runs of equal group -> one launch each   (flags + scans)
//   5. work items {ket run, first T occurrence, count <= bras per item}                                (list_items_kernel)
// and the host only reads back the (small) group table to launch one kernel per class pair.
// The reference has no batched call; its callers loop over quartets from OpenMP threads (examples/time_c60.c:196-219).
#include <cstdio>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include "../../include/cint_b200.h"
#include "driver.h"
#include "engine.h"
#include "rys.cuh"

int rys_tab_off(int nroots);

#define CU_OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return b200_fail(CINTB200_ENODEV, "%s failed: %s", #call, cudaGetErrorString(e_)); } while (0)

struct ListDevArgs {
    const int *shls; long long n; int ncenter, nbas, ncls, cart;
    const PairHdr *pairs; const int *cls_of, *row_of, *sdim, *per;     // sdim[2][nbas] (sph, cart); per[2][ncls*ncls]
    unsigned long long *key; unsigned int *val;
    long long *size;                    // block size per tuple (input of the offset scan)
    int *tsel, *sa, *sb, *sc, *sd, *ket, *nz;
    int *counters;                      // [0] tuples with a bad shell id, [1] tuples without a specialised kernel
};

__global__ void list_key_kernel(const ListDevArgs A)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= A.n) return;
    const int *s = A.shls + t * A.ncenter;
    const long long npair2 = (long long)A.nbas * (A.nbas + 1) / 2;
    int sh[4] = {s[0], s[1], A.ncenter > 2 ? s[2] : 0, A.ncenter > 3 ? s[3] : 0};
    bool bad = false;
    for (int m = 0; m < A.ncenter; m++) bad |= sh[m] < 0 || sh[m] >= A.nbas;
    if (bad) {
        atomicAdd(A.counters, 1);
        A.key[t] = ~0ull; A.val[t] = (unsigned int)t; A.size[t] = 0; A.nz[t] = 0;
        return;
    }
    const int *dim = A.sdim + (A.cart ? A.nbas : 0);
    int bra, ket, sa, sb;
    long long sc, sd, size;
    if (A.ncenter == 2) {               // (i|k): single-shell pseudo pairs on both sides
        const int i = sh[0], k = sh[1];
        bra = (int)(npair2 + i); ket = (int)(npair2 + k);
        sa = 1; sb = 0; sc = dim[i]; sd = 0;
        size = (long long)dim[i] * dim[k];
    } else {
        const int i = sh[0], j = sh[1], k = sh[2], l = sh[3];
        const long long di = dim[i], dj = dim[j], dk = dim[k], dl = A.ncenter == 4 ? dim[l] : 1;
        bra = (int)(i >= j ? (long long)i * (i + 1) / 2 + j : (long long)j * (j + 1) / 2 + i);
        const bool a_is_i = A.pairs[bra].sh_a == i;
        sa = a_is_i ? 1 : (int)di; sb = a_is_i ? (int)di : 1;
        if (A.ncenter == 4) {
            ket = (int)(k >= l ? (long long)k * (k + 1) / 2 + l : (long long)l * (l + 1) / 2 + k);
            const bool c_is_k = A.pairs[ket].sh_a == k;
            sc = c_is_k ? di * dj : di * dj * dk; sd = c_is_k ? di * dj * dk : di * dj;
        } else {
            ket = (int)(npair2 + k);
            sc = di * dj; sd = 0;
        }
        size = di * dj * dk * dl;
    }
    const int g = A.cls_of[bra] * A.ncls + A.cls_of[ket];
    if (A.per[(A.cart ? A.ncls * A.ncls : 0) + g] == 0) atomicAdd(A.counters + 1, 1);
    const int nppb = A.pairs[bra].npp, nppk = A.pairs[ket].npp;
    A.key[t] = ((unsigned long long)g << 42) | ((unsigned long long)(unsigned int)ket << 8) | ((unsigned long long)(sc > sd) << 7)
               | (unsigned long long)(127 - min(nppb, 127));
    A.val[t] = (unsigned int)t;
    A.size[t] = size;
    A.tsel[t] = A.row_of[bra];
    A.sa[t] = sa; A.sb[t] = sb; A.sc[t] = (int)sc; A.sd[t] = (int)sd; A.ket[t] = ket;
    A.nz[t] = nppb > 0 && nppk > 0;
}

// sorted position q -> flags: starts a ket run / starts a group
__global__ void list_flag_kernel(const unsigned long long *__restrict__ key, long long m, int *runflag, int *grpflag)
{
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    const unsigned long long k = key[q], p = q ? key[q - 1] : ~k;
    runflag[q] = q == 0 || (k >> 7) != (p >> 7);
    grpflag[q] = q == 0 || (k >> 42) != (p >> 42);
}

// per sorted position: gather the tuple's arrays into sorted order; run / group starts record their positions
__global__ void list_gather_kernel(const unsigned long long *__restrict__ key, const unsigned int *__restrict__ val, long long m, long long nruns_cap,
                                   const int *__restrict__ runflag, const int *__restrict__ runid, const int *__restrict__ grpflag,
                                   const int *__restrict__ grpid, const long long *__restrict__ offs, const int *__restrict__ tsel,
                                   const int *__restrict__ sa, const int *__restrict__ sb, const int *__restrict__ sc, const int *__restrict__ sd,
                                   const int *__restrict__ ket, int *o_tsel, long long *o_trow, int *o_tstride, int *o_upair, int *o_ustride,
                                   long long *run_start, int4 *groups)
{
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    const unsigned int t = val[q];
    o_tsel[q] = tsel[t];
    o_trow[q] = offs[t];
    o_tstride[q] = sa[t]; o_tstride[m + q] = sb[t];
    if (runflag[q]) {
        const int r = runid[q];                     // inclusive scan - 1 done by the caller (exclusive scan of flags, + flag)
        o_upair[r] = ket[t];
        o_ustride[r] = sc[t]; o_ustride[nruns_cap + r] = sd[t];
        run_start[r] = q;
    }
    if (grpflag[q]) groups[grpid[q]] = make_int4((int)(key[q] >> 42), (int)q, runid[q], 0);    // {group key, first position, first run, -}
}

// per run: number of work items (bras per item depends on the group's kernel)
__global__ void list_runitems_kernel(const unsigned long long *__restrict__ key, const long long *__restrict__ run_start, long long nruns, long long m,
                                     const int *__restrict__ per, int *nitems)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nruns) return;
    const long long q0 = run_start[r], q1 = r + 1 < nruns ? run_start[r + 1] : m;
    const int p = per[(int)(key[q0] >> 42)];
    nitems[r] = p > 0 ? (int)((q1 - q0 + p - 1) / p) : 0;
}

// per run: its work items {ket run local to the group, first T occurrence local to the group, count, -}
__global__ void list_items_kernel(const unsigned long long *__restrict__ key, const long long *__restrict__ run_start, long long nruns, long long m,
                                  const int *__restrict__ per, const int *__restrict__ itemoff, const int *__restrict__ grp_of_run,
                                  const int4 *__restrict__ groups, int4 *items)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nruns) return;
    const long long q0 = run_start[r], q1 = r + 1 < nruns ? run_start[r + 1] : m;
    const int p = per[(int)(key[q0] >> 42)];
    if (p <= 0) return;
    const int4 G = groups[grp_of_run[r]];
    int4 *it = items + itemoff[r];
    int k = 0;
    for (long long q = q0; q < q1; q += p, k++)
        it[k] = make_int4((int)(r - G.z), (int)(q - G.y), (int)min((long long)p, q1 - q), 0);
}

// group of a run: groups started up to and including the run's first position, minus one (grpid is an EXCLUSIVE scan)
__global__ void list_grp_of_run_kernel(const long long *__restrict__ run_start, long long nruns, const int *__restrict__ grpid,
                                       const int *__restrict__ grpflag, int *grp_of_run)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nruns) grp_of_run[r] = grpid[run_start[r]] + grpflag[run_start[r]] - 1;
}

static int listdev_tables(CINTOpt *c)
{
    ListTables *lt = c->ltab;
    if (lt->d_cls_of) return 0;
    const int ncls = (int)lt->cls.size();
    std::vector<int> sdim(2 * (size_t)c->nbas), per(2 * (size_t)ncls * ncls, 0);
    for (int i = 0; i < c->nbas; i++) {
        sdim[i] = (2 * c->shells[i].l + 1) * c->shells[i].nctr;
        sdim[c->nbas + i] = B200_NCART(c->shells[i].l) * c->shells[i].nctr;
    }
    for (int cart = 0; cart < 2; cart++)
        for (int g = 0; g < ncls * ncls; g++) {
            const ListChoice &ch = (cart ? lt->choice_cart : lt->choice)[g];
            per[(size_t)cart * ncls * ncls + g] = ch.fn ? (ch.coop ? 32 / ch.ci.fs : 32) : 0;
        }
    auto up = [&](int **dst, const std::vector<int> &v) {
        return b200_dmalloc((void **)dst, sizeof(int) * std::max<size_t>(1, v.size())) == cudaSuccess &&
               cudaMemcpy(*dst, v.data(), sizeof(int) * v.size(), cudaMemcpyHostToDevice) == cudaSuccess;
    };
    if (!up(&lt->d_cls_of, lt->cls_of) || !up(&lt->d_row_of, lt->row_of) || !up(&lt->d_sdim, sdim) || !up(&lt->d_per, per))
        return b200_fail(CINTB200_ENOMEM, "list mode: device tables");
    return 0;
}

// Returns 1 when the whole list was evaluated by the specialised kernels (launches queued on c->stream, not synchronised;
// *total = elements of the packed output / highest offset reached), 0 when the caller has to use the host path (a class without
// a specialised kernel, output beyond 2^31 elements, ...), negative on error.  out_off: optional host offsets.  Caller holds c->mtx.
int list_mode_device(CINTOpt *c, int ncenter, int cart, const int *shls, size_t n, const size_t *out_off, double *d_out_or_null,
                     double **d_out_used, size_t *total, int *nonzero)
{
    double tph = b200_now();
    if (!c->ltab && listtables_build(c)) return CINTB200_ENOMEM;
    ListTables *lt = c->ltab;
    if (listdev_tables(c)) return CINTB200_ENOMEM;
    b200_phase("list: tables", tph); tph = b200_now();
    const int ncls = (int)lt->cls.size();
    if ((unsigned long long)ncls * ncls >= (1ull << 22) || n >= (1ull << 31)) return 0;
    cudaStream_t st = c->stream;
    // ---- carve the work area ----
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = al(off + bytes); return o; };
    const size_t o_shls = take(sizeof(int) * n * ncenter), o_key = take(8 * n), o_key2 = take(8 * n), o_val = take(4 * n), o_val2 = take(4 * n);
    const size_t o_size = take(8 * n), o_offs = take(8 * n), o_tsel = take(4 * n), o_sa = take(4 * n), o_sb = take(4 * n), o_sc = take(4 * n);
    const size_t o_sd = take(4 * n), o_ket = take(4 * n), o_nz = take(4 * n), o_cnt = take(64);
    const size_t o_rf = take(4 * n), o_ri = take(4 * n), o_gf = take(4 * n), o_gi = take(4 * n);
    const size_t o_otsel = take(4 * n), o_otrow = take(8 * n), o_otstr = take(8 * n), o_oupair = take(4 * n), o_oustr = take(8 * n);
    const size_t o_rstart = take(8 * n), o_groups = take(sizeof(int4) * std::min<size_t>(n, (size_t)ncls * ncls) + 64), o_nit = take(4 * n), o_itoff = take(4 * n + 4);
    const size_t o_gor = take(4 * n), o_ucol = take(8 * n), o_lcnt = take(4 * std::min<size_t>(n, (size_t)ncls * ncls) + 64);
    size_t tmp_bytes = 0, tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (unsigned int *)nullptr, (unsigned int *)nullptr, (int)n, 0, 64, st);
    tmp_bytes = std::max(tmp_bytes, tb);
    cub::DeviceScan::ExclusiveSum(nullptr, tb, (long long *)nullptr, (long long *)nullptr, (int)n, st);
    tmp_bytes = std::max(tmp_bytes, tb);
    cub::DeviceScan::ExclusiveSum(nullptr, tb, (int *)nullptr, (int *)nullptr, (int)n + 1, st);
    tmp_bytes = std::max(tmp_bytes, tb);
    const size_t o_tmp = take(tmp_bytes);
    // items: at most one per tuple plus one per run
    const size_t o_items = take(sizeof(int4) * 2 * n);
    if (lt->cap_work < off) {
        cudaStreamSynchronize(st);
        b200_big_free(lt->d_work); lt->d_work = nullptr; lt->cap_work = 0;
        if (b200_big_alloc(&lt->d_work, off + off / 4)) return b200_fail(CINTB200_ENOMEM, "list mode: %zu bytes of device work area", off);
        lt->cap_work = off + off / 4;
    }
    char *w = (char *)lt->d_work;
    ListDevArgs A;
    A.shls = (const int *)(w + o_shls); A.n = (long long)n; A.ncenter = ncenter; A.nbas = c->nbas; A.ncls = ncls; A.cart = cart;
    A.pairs = c->d_pairs; A.cls_of = lt->d_cls_of; A.row_of = lt->d_row_of; A.sdim = lt->d_sdim; A.per = lt->d_per;
    A.key = (unsigned long long *)(w + o_key); A.val = (unsigned int *)(w + o_val); A.size = (long long *)(w + o_size);
    A.tsel = (int *)(w + o_tsel); A.sa = (int *)(w + o_sa); A.sb = (int *)(w + o_sb); A.sc = (int *)(w + o_sc); A.sd = (int *)(w + o_sd);
    A.ket = (int *)(w + o_ket); A.nz = (int *)(w + o_nz); A.counters = (int *)(w + o_cnt);
    const int *per = lt->d_per + (cart ? ncls * ncls : 0);
    const unsigned nb = (unsigned)((n + 255) / 256);
    CU_OK(cudaMemcpyAsync(w + o_shls, shls, sizeof(int) * n * ncenter, cudaMemcpyHostToDevice, st));
    CU_OK(cudaMemsetAsync(w + o_cnt, 0, 64, st));
    list_key_kernel<<<nb, 256, 0, st>>>(A);
    long long *d_offs = (long long *)(w + o_offs);
    if (out_off) CU_OK(cudaMemcpyAsync(d_offs, out_off, 8 * n, cudaMemcpyHostToDevice, st));
    else { tb = tmp_bytes; cub::DeviceScan::ExclusiveSum(w + o_tmp, tb, A.size, d_offs, (int)n, st); }
    tb = tmp_bytes;
    cub::DeviceRadixSort::SortPairs(w + o_tmp, tb, A.key, (unsigned long long *)(w + o_key2), A.val, (unsigned int *)(w + o_val2), (int)n, 0, 64, st);
    const unsigned long long *skey = (const unsigned long long *)(w + o_key2);
    const unsigned int *sval = (const unsigned int *)(w + o_val2);
    int *runflag = (int *)(w + o_rf), *runid = (int *)(w + o_ri), *grpflag = (int *)(w + o_gf), *grpid = (int *)(w + o_gi);
    list_flag_kernel<<<nb, 256, 0, st>>>(skey, (long long)n, runflag, grpflag);
    // exclusive scans of the flags: id of position q = (number of starts before q) + flag - 1 = exclusive + flag - 1; using the
    // inclusive form through an exclusive scan over n + 1 entries shifted by one is not needed: ids are taken at the starts only,
    // where exclusive(q) is exactly the index of the run / group that starts at q
    tb = tmp_bytes; cub::DeviceScan::ExclusiveSum(w + o_tmp, tb, runflag, runid, (int)n, st);
    tb = tmp_bytes; cub::DeviceScan::ExclusiveSum(w + o_tmp, tb, grpflag, grpid, (int)n, st);
    // counters + totals needed on the host: bad ids, unhandled, number of runs and groups, last offset + size
    struct Tail { int cnt[2]; int lastrunflag, lastrunid, lastgrpflag, lastgrpid; long long lastoff, lastsize; } tail;
    CU_OK(cudaMemcpyAsync(tail.cnt, w + o_cnt, 8, cudaMemcpyDeviceToHost, st));
    CU_OK(cudaMemcpyAsync(&tail.lastrunflag, runflag + n - 1, 4, cudaMemcpyDeviceToHost, st));
    CU_OK(cudaMemcpyAsync(&tail.lastrunid, runid + n - 1, 4, cudaMemcpyDeviceToHost, st));
    CU_OK(cudaMemcpyAsync(&tail.lastgrpflag, grpflag + n - 1, 4, cudaMemcpyDeviceToHost, st));
    CU_OK(cudaMemcpyAsync(&tail.lastgrpid, grpid + n - 1, 4, cudaMemcpyDeviceToHost, st));
    CU_OK(cudaMemcpyAsync(&tail.lastoff, d_offs + n - 1, 8, cudaMemcpyDeviceToHost, st));
    CU_OK(cudaMemcpyAsync(&tail.lastsize, A.size + n - 1, 8, cudaMemcpyDeviceToHost, st));
    CU_OK(cudaStreamSynchronize(st));
    b200_phase("list: keys + sort + flags", tph); tph = b200_now();
    if (tail.cnt[0]) return b200_fail(CINTB200_EINVAL, "%d tuples with a shell id out of range", tail.cnt[0]);
    if (tail.cnt[1]) return 0;                          // some class has no specialised kernel: host path (generic kernel for those)
    size_t tot = 0;
    if (!out_off) tot = (size_t)(tail.lastoff + tail.lastsize);
    else {
        // highest end of a block: offsets are the caller's, sizes ours -- take the maximum on the host from the block sizes
        std::vector<long long> hs(n);
        CU_OK(cudaMemcpy(hs.data(), A.size, 8 * n, cudaMemcpyDeviceToHost));
        tot = 0;
        for (size_t t = 0; t < n; t++) tot = std::max(tot, out_off[t] + (size_t)hs[t]);
    }
    if (tot >= ((size_t)1 << 31)) return 0;             // the kernels keep row offsets in 32 bits
    *total = tot;
    double *d_out = d_out_or_null;
    if (!d_out) {
        if (ctx_reserve(c, (void **)&c->d_out, &c->cap_out, sizeof(double) * std::max<size_t>(1, tot), false)) return CINTB200_ENOMEM;
        d_out = c->d_out;
    }
    *d_out_used = d_out;
    const long long nruns = tail.lastrunid + tail.lastrunflag, ngroups = tail.lastgrpid + tail.lastgrpflag;    // exclusive scans + last flag
    int4 *groups = (int4 *)(w + o_groups);
    long long *run_start = (long long *)(w + o_rstart);
    list_gather_kernel<<<nb, 256, 0, st>>>(skey, sval, (long long)n, (long long)n, runflag, runid, grpflag, grpid, d_offs, A.tsel, A.sa, A.sb, A.sc, A.sd, A.ket,
                                           (int *)(w + o_otsel), (long long *)(w + o_otrow), (int *)(w + o_otstr), (int *)(w + o_oupair),
                                           (int *)(w + o_oustr), run_start, groups);
    const unsigned nbr = (unsigned)((nruns + 255) / 256);
    int *nitems = (int *)(w + o_nit), *itemoff = (int *)(w + o_itoff), *grp_of_run = (int *)(w + o_gor);
    list_runitems_kernel<<<nbr, 256, 0, st>>>(skey, run_start, nruns, (long long)n, per, nitems);
    CU_OK(cudaMemsetAsync(nitems + nruns, 0, 4, st));
    tb = tmp_bytes; cub::DeviceScan::ExclusiveSum(w + o_tmp, tb, nitems, itemoff, (int)nruns + 1, st);
    list_grp_of_run_kernel<<<nbr, 256, 0, st>>>(run_start, nruns, grpid, grpflag, grp_of_run);
    list_items_kernel<<<nbr, 256, 0, st>>>(skey, run_start, nruns, (long long)n, per, itemoff, grp_of_run, groups, (int4 *)(w + o_items));
    CU_OK(cudaMemsetAsync(w + o_ucol, 0, 8 * (size_t)nruns, st));
    CU_OK(cudaMemsetAsync(w + o_lcnt, 0, 4 * (size_t)ngroups, st));
    // group table + item offsets of the group starts to the host
    std::vector<int4> hg((size_t)ngroups);
    CU_OK(cudaMemcpyAsync(hg.data(), groups, sizeof(int4) * ngroups, cudaMemcpyDeviceToHost, st));
    std::vector<int> h_itoff((size_t)nruns + 1);
    CU_OK(cudaMemcpyAsync(h_itoff.data(), itemoff, 4 * ((size_t)nruns + 1), cudaMemcpyDeviceToHost, st));
    CU_OK(cudaStreamSynchronize(st));
    b200_phase("list: runs + items", tph); tph = b200_now();
    for (long long g = 0; g < ngroups; g++) {
        const int gkey = hg[g].x;
        const long long t0 = hg[g].y, u0 = hg[g].z;
        const long long t1 = g + 1 < ngroups ? hg[g + 1].y : (long long)n, u1 = g + 1 < ngroups ? hg[g + 1].z : nruns;
        const ListChoice &ch = (cart ? lt->choice_cart : lt->choice)[gkey];
        if (listclass_upload(c, lt->cls[gkey / ncls])) return CINTB200_ENOMEM;
        const ListClass &B = lt->cls[gkey / ncls], &K = lt->cls[gkey % ncls];
        const int nroots = (B.la + B.lb + K.la + K.lb) / 2 + 1;
        TileParams P;
        memset(&P, 0, sizeof P);
        P.tprim = B.d_tprim; P.tgeom = B.d_tgeom; P.tnpp = B.d_tnpp;
        P.NT = (int)B.ids.size(); P.Q = B.Q;
        P.trow = (const long long *)(w + o_otrow) + t0;
        P.tstride = (const int *)(w + o_otstr) + t0;          // rows of the [2][n] array: second row n entries further
        P.NTs = (int)n;
        P.tsel = (const int *)(w + o_otsel) + t0;
        P.t_begin = 0; P.t_end = (int)(t1 - t0); P.nca_t = B.nca;
        P.upair = (const int *)(w + o_oupair) + u0;
        P.ucol = (const long long *)(w + o_ucol);
        P.ustride = (const int *)(w + o_oustr) + u0;
        P.NU = (int)(u1 - u0); P.NU_all = (int)n; P.u_step = 1; P.u_first = 0; P.nca_u = K.nca; P.umax = std::max(1, K.Q);
        P.out = d_out; P.row0 = 0; P.ld = 1;
        P.pairs = c->d_pairs; P.prims = c->d_prims; P.pcoef = c->d_pcoef;
        P.rys = (!ch.coop && REG_FAST_RYS && nroots <= RYS_FNMAX && nroots <= REG_FAST_NMAX) ? c->d_rys_fast + rys_fast_off(nroots)
                                                                                          : c->d_rys + rys_tab_off(nroots);
        P.rs_w2 = c->omega * c->omega; P.rs_sign = c->omega; P.rs_pass0 = c->omega > 0 ? 1 : 0;
        P.items = (const int4 *)(w + o_items) + h_itoff[u0];
        P.nitems = (long long)(h_itoff[u1] - h_itoff[u0]);
        P.gx = 1;
        P.counter = (unsigned int *)(w + o_lcnt) + g;
        P.batch = (int)std::max<long long>(1, std::min<long long>(16, P.nitems / (148 * 32 * 4)));
        if (P.nitems == 0) continue;
        if (ch.coop ? coop_kernel_launch(ch.fn, ch.ci, K.nca * K.ncb, P, 1, 1, st) : reg_kernel_launch(ch.fn, nroots, K.nca * K.ncb, P, 1, 1, st))
            return b200_fail(CINTB200_ENODEV, "list-mode kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        c->launches++;
    }
    if (nonzero) CU_OK(cudaMemcpyAsync(nonzero, A.nz, 4 * n, cudaMemcpyDeviceToHost, st));
    b200_phase("list: launches queued", tph);
    return 1;
}
