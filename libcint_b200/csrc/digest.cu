// Consumers of the whole-job tiles (driver.cu): everything that reads a finished tile while it is still in HBM.
//
//  * per-row CHECKSUMS of the valid entries (sum v, sum |v|, sum v * g(c,d)), reduced on the host to per-bra-pair
//    fingerprints that do not depend on chunking or on the rank sharding -- the whole-job parity check against
//    oracle/ref_golden.c (examples/time_c60.c:200-219 discards `buf`; this is what lets every block be value-checked);
//  * J/K DIGESTION: Coulomb and exchange matrices from the 8-fold unique quartets, tile by tile, so that no
//    integral has to leave the GPU.  The reference has no such routine (its callers -- pyscf's CVHFnr_* -- digest the
//    per-quartet `buf` of int2e_sph); the statement it is checked against is the plain 8-image loop of oracle/ref_golden.c.
//
// A tile is column-major T[row(ab) + ld * col(cd)]; rows = AO pairs of the chunk's bra shell pairs, columns = AO pairs of
// this rank's kets, both laid out like the reference's per-quartet buf (first index fastest).  Entries with K > I are never
// written by the ERI kernels; entries with K == I, L > J are written but redundant (the loop of time_c60.c evaluates them,
// the 8-fold unique set does not contain them).  Every consumer masks with kl <= ij (pair order) or K <= I.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include "../../include/cint_b200.h"
#include "driver.h"
#include "engine.h"

#define CU_OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return b200_fail(CINTB200_ENODEV, "%s failed: %s", #call, cudaGetErrorString(e_)); } while (0)

struct JKUnit { int ao0, nx, ebeg, eend; };          // AOs [ao0, ao0 + nx) of one shell ("outer" index x) and its visit list
struct JKEntry { int colbase, kl, aoY, info; };      // info = dy | dk << 8 | xk << 16;  col = colbase + x * sx + y * sy
#define JK_NXMAX 5

struct DigestState {
    long long nrows = 0, ncols = 0, ldmax = 0;
    int nao = 0;
    // checksums
    int *d_rowI = nullptr, *d_colK = nullptr;
    double *d_colg = nullptr, *d_rowsums = nullptr;  // [3][nrows]
    int have_rowsums = 0;
    // J/K
    int jk_ready = 0;
    int4 *d_rowinfo = nullptr;                       // {a, b, ij, I == J}
    int2 *d_colinfo = nullptr;                       // {kl, K == L}
    int *d_colc = nullptr, *d_cold = nullptr;
    JKUnit *d_units = nullptr; JKEntry *d_entries = nullptr;
    int unit_beg[JK_NXMAX + 2] = {0};                // units sorted by nx: [unit_beg[nx], unit_beg[nx + 1])
    double *d_dm = nullptr, *d_Dab = nullptr, *d_Dcd = nullptr;
    double *d_PA = nullptr, *d_PB = nullptr, *d_Jp = nullptr, *d_Kp = nullptr, *d_jrow = nullptr, *d_jcol = nullptr;
};

void digest_free(DigestState *d)
{
    if (!d) return;
    b200_dfree(d->d_rowI); b200_dfree(d->d_colK); b200_dfree(d->d_colg); b200_dfree(d->d_rowsums);
    b200_dfree(d->d_rowinfo); b200_dfree(d->d_colinfo); b200_dfree(d->d_colc); b200_dfree(d->d_cold); b200_dfree(d->d_units); b200_dfree(d->d_entries);
    b200_dfree(d->d_dm); b200_dfree(d->d_Dab); b200_dfree(d->d_Dcd); b200_big_free(d->d_PA); b200_big_free(d->d_PB);
    b200_dfree(d->d_Jp); b200_dfree(d->d_Kp); b200_dfree(d->d_jrow); b200_dfree(d->d_jcol);
    delete d;
}

template <class T>
static int up(T **dst, const std::vector<T> &src)
{
    if (b200_dmalloc((void **)dst, sizeof(T) * std::max<size_t>(1, src.size())) != cudaSuccess)
        return b200_fail(CINTB200_ENOMEM, "cudaMalloc of %zu bytes failed", sizeof(T) * src.size());
    if (!src.empty() && cudaMemcpy(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice) != cudaSuccess)
        return b200_fail(CINTB200_ENODEV, "upload failed");
    return 0;
}

static inline int shell_dim_of(const ShellInfo &s, int cart) { return (cart ? B200_NCART(s.l) : 2 * s.l + 1) * s.nctr; }
static inline int shell_ao_of(const ShellInfo &s, int cart) { return cart ? s.ao_cart : s.ao_sph; }
static inline void pair_shells(int p, int *i, int *j)
{
    int ii = (int)((sqrt(8.0 * p + 1.0) - 1.0) / 2.0);
    while ((long long)(ii + 1) * (ii + 2) / 2 <= p) ii++;
    while ((long long)ii * (ii + 1) / 2 > p) ii--;
    *i = ii; *j = p - ii * (ii + 1) / 2;
}

// ------------------------------------------------------------------ checksums
// weights of oracle/ref_golden.c
static inline double g_weight(int c, int d) { return cos(0.37 * c + 0.61 * d + 0.5); }
static inline double h_weight(int r) { return cos(0.91 * r + 0.3); }

__global__ void __launch_bounds__(256) tile_rowsum_kernel(const double *__restrict__ tile, long long ld, long long ncols, long long row0,
                                                           const int *__restrict__ rowI, const int *__restrict__ colK,
                                                           const double *__restrict__ colg, double *rowsums, long long nrows_total,
                                                           int cols_per_block)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= ld) return;
    const int I = rowI[row0 + r];
    const long long c0 = (long long)blockIdx.y * cols_per_block;
    const long long c1 = c0 + cols_per_block < ncols ? c0 + cols_per_block : ncols;
    double s[2] = {0, 0}, a[2] = {0, 0}, f[2] = {0, 0};
    const double *t = tile + r;
    long long c = c0;
    for (; c + 1 < c1; c += 2) {            // two independent chains; entries with K > I were never written: do not read them
        const bool v0 = colK[c] <= I, v1 = colK[c + 1] <= I;
        const double x0 = v0 ? t[ld * c] : 0.0, x1 = v1 ? t[ld * (c + 1)] : 0.0;
        s[0] += x0; a[0] += fabs(x0); f[0] = fma(x0, colg[c], f[0]);
        s[1] += x1; a[1] += fabs(x1); f[1] = fma(x1, colg[c + 1], f[1]);
    }
    if (c < c1 && colK[c] <= I) { const double x0 = t[ld * c]; s[0] += x0; a[0] += fabs(x0); f[0] = fma(x0, colg[c], f[0]); }
    atomicAdd(rowsums + row0 + r, s[0] + s[1]);
    atomicAdd(rowsums + nrows_total + row0 + r, a[0] + a[1]);
    atomicAdd(rowsums + 2 * nrows_total + row0 + r, f[0] + f[1]);
}

static int checksums_prepare(CINTOpt *c, JobPlan *plan, DigestState *d)
{
    if (d->d_rowI) return 0;
    std::vector<int> rowI(d->nrows), colK(d->ncols);
    std::vector<double> colg(d->ncols);
    for (long long r = 0; r < d->nrows; r++) { int i, j; pair_shells(plan->row_pair[r], &i, &j); rowI[r] = i; }
    const int cart = plan->cart;
    const int ao_aux0 = plan->ncenter == 3 ? shell_ao_of(c->shells[plan->aux0], cart) : 0;
    for (long long q = 0; q < d->ncols; q++) {
        if (plan->ncenter == 3) {           // columns = auxiliary functions: always valid; weight g(c - first auxiliary AO, 0)
            const ShellInfo &sk = c->shells[plan->col_pair[q]];
            colK[q] = -1;
            colg[q] = g_weight(shell_ao_of(sk, cart) + plan->col_pos[q] - ao_aux0, 0);
        } else {
            int k, l;
            pair_shells(plan->col_pair[q], &k, &l);
            const int dk = shell_dim_of(c->shells[k], cart);
            colK[q] = k;
            colg[q] = g_weight(shell_ao_of(c->shells[k], cart) + plan->col_pos[q] % dk, shell_ao_of(c->shells[l], cart) + plan->col_pos[q] / dk);
        }
    }
    if (up(&d->d_rowI, rowI) || up(&d->d_colK, colK) || up(&d->d_colg, colg)) return CINTB200_ENOMEM;
    if (b200_dmalloc((void **)&d->d_rowsums, sizeof(double) * 3 * std::max<long long>(1, d->nrows)) != cudaSuccess)
        return b200_fail(CINTB200_ENOMEM, "cannot allocate the row-sum buffer");
    return 0;
}

// ------------------------------------------------------------------ J/K digestion
struct JKArgs {
    const double *tile; long long ld, row0; int nao;
    const int4 *rowinfo; const JKUnit *units; const JKEntry *entries; int ubeg, uend;
    const double *dm, *Dcd;                 // dm row-major nao x nao (symmetric); Dcd[col] = 2 f_KL D[c,d]
    double *PA, *PB, *jrow; long long ldP;
    int want_k;
};

// One thread per tile row (ab): for every outer unit X (AOs x of one ket-side shell) it walks the kets that contain X and
// accumulates  kA[x] = sum_y (ab|xy) D[b,y]  and  kB[x] = sum_y (ab|xy) D[a,y]  in registers (y = the ket's other index),
// i.e. its share of K'[a,x] and K'[b,x]; each ket pair is visited from both of its shells, so only this one update type
// is needed and nothing is scattered.  Tile loads are coalesced (consecutive threads = consecutive rows of one column).
// The partial sums go to private slots PA / PB[x][row] and are folded over the rows sharing a (or b) by jk_fold_kernel.
// The Coulomb row part J'[a,b] += 2 s (ab|cd) D[c,d] rides along in the visit from the ket's first shell.
// JK_RPT tile rows per thread: more loads in flight per thread (ncu: the kernel waits on HBM latency with 16 resident warps per SM),
// the per-entry decoding shared by all of them.  Measured alternative (round 2, removed): the column segments staged by per-warp
// rings of cp.async.bulk copies completing on mbarriers (16 segments in flight per warp, density values prefetched one item
// ahead) -- 1516 vs 1502 ms for the C60 J/K pass: so the plateau at ~3.5 TB/s is not the SM-side load
// latency the stall counters point at (not understood further; candidates: 2 KB column segments one leading dimension apart,
// the dependent density gathers).
template <int NX, int JK_RPT>
__global__ void __launch_bounds__(128) jk_rows_kernel(const JKArgs A)
{
    const long long row0 = (long long)blockIdx.x * (128 * JK_RPT) + threadIdx.x;
    bool active[JK_RPT];
    int a[JK_RPT], b[JK_RPT], ij[JK_RPT];
    double fij[JK_RPT], jr[JK_RPT];
    const double *trow[JK_RPT];
    int maxij = -1;
#pragma unroll
    for (int q = 0; q < JK_RPT; q++) {
        const long long row = row0 + 128 * q;
        active[q] = row < A.ld;
        const int4 ri = active[q] ? A.rowinfo[A.row0 + row] : make_int4(0, 0, -1, 0);
        a[q] = ri.x; b[q] = ri.y; ij[q] = ri.z; fij[q] = ri.w ? 0.5 : 1.0; jr[q] = 0.0;
        trow[q] = A.tile + (active[q] ? row : 0);
        maxij = max(maxij, ij[q]);
    }
    maxij = __reduce_max_sync(0xffffffffu, maxij);
    for (int u = A.ubeg + blockIdx.y; u < A.uend; u += gridDim.y) {
        const JKUnit un = A.units[u];
        double kA[JK_RPT][NX], kB[JK_RPT][NX];
#pragma unroll
        for (int q = 0; q < JK_RPT; q++)
#pragma unroll
            for (int x = 0; x < NX; x++) kA[q][x] = kB[q][x] = 0.0;
        for (int e = un.ebeg; e < un.eend; e++) {
            const JKEntry en = A.entries[e];
            if (en.kl > maxij) break;                       // entries are sorted by kl: nothing further is valid for this warp
            bool valid[JK_RPT];
            double w[JK_RPT];
#pragma unroll
            for (int q = 0; q < JK_RPT; q++) {
                valid[q] = en.kl <= ij[q];
                w[q] = valid[q] ? (en.kl == ij[q] ? 0.5 * fij[q] : fij[q]) : 0.0;
            }
            const int dy = en.info & 255, dk = (en.info >> 8) & 255, xk = en.info >> 16;
            const long long sx = (xk ? 1 : dk) * A.ld, sy = (xk ? dk : 1) * A.ld;
            const long long o0 = (long long)en.colbase * A.ld;
            for (int y = 0; y < dy; y++) {
                const double *Dy = A.dm + (size_t)(en.aoY + y) * A.nao;
                double v[JK_RPT][NX];
#pragma unroll
                for (int q = 0; q < JK_RPT; q++)
#pragma unroll
                    for (int x = 0; x < NX; x++) v[q][x] = valid[q] ? trow[q][o0 + y * sy + x * sx] : 0.0;   // never read unwritten entries
#pragma unroll
                for (int q = 0; q < JK_RPT; q++) {
                    const double dA = w[q] * Dy[a[q]], dB = w[q] * Dy[b[q]];
#pragma unroll
                    for (int x = 0; x < NX; x++) {
                        kA[q][x] = fma(v[q][x], dB, kA[q][x]);
                        kB[q][x] = fma(v[q][x], dA, kB[q][x]);
                        if (xk) jr[q] = fma(v[q][x] * w[q], A.Dcd[en.colbase + x + y * dk], jr[q]);
                    }
                }
            }
        }
        if (A.want_k) {
#pragma unroll
            for (int q = 0; q < JK_RPT; q++)
                if (active[q]) {
#pragma unroll
                    for (int x = 0; x < NX; x++) {
                        A.PA[(size_t)(un.ao0 + x) * A.ldP + row0 + 128 * q] = kA[q][x];
                        A.PB[(size_t)(un.ao0 + x) * A.ldP + row0 + 128 * q] = kB[q][x];
                    }
                }
        }
    }
#pragma unroll
    for (int q = 0; q < JK_RPT; q++)
        if (active[q] && jr[q] != 0.0) atomicAdd(A.jrow + A.row0 + row0 + 128 * q, jr[q]);
}

// K'[a, x] += sum over the chunk's rows with first AO a of PA[x][row];  K'[b, x] += ... PB[x][row].  One block per x.
__global__ void __launch_bounds__(256) jk_fold_kernel(const double *__restrict__ PA, const double *__restrict__ PB, long long ldP, long long ld,
                                                       long long row0, const int4 *__restrict__ rowinfo, int nao, double *Kp)
{
    extern __shared__ double acc[];
    const int x = blockIdx.x;
    for (int i = threadIdx.x; i < nao; i += blockDim.x) acc[i] = 0.0;
    __syncthreads();
    for (long long r = threadIdx.x; r < ld; r += blockDim.x) {
        const int4 ri = rowinfo[row0 + r];
        const double va = PA[(size_t)x * ldP + r], vb = PB[(size_t)x * ldP + r];
        if (va != 0.0) atomicAdd(acc + ri.x, va);
        if (vb != 0.0) atomicAdd(acc + ri.y, vb);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nao; i += blockDim.x)
        if (acc[i] != 0.0) Kp[(size_t)i * nao + x] += acc[i];      // this block is the only writer of column x; launches are stream-ordered
}

// Coulomb column part: J'[c,d] += 2 s sum_ab (ab|cd) D[a,b].  One warp per column, lanes stride over the rows (coalesced).
__global__ void __launch_bounds__(256) jk_cols_kernel(const double *__restrict__ tile, long long ld, long long ncols, long long row0,
                                                       const int4 *__restrict__ rowinfo, const int2 *__restrict__ colinfo,
                                                       const double *__restrict__ Dab, double *jcol)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long col = warp; col < ncols; col += nwarp) {
        const int2 ci = colinfo[col];
        const double *t = tile + ld * col;
        double p0 = 0.0, p1 = 0.0;
        long long r = lane;
        for (; r + 32 < ld; r += 64) {              // two independent chains (four were measured slower: 144 vs 107 ms per pass)
            const int ij0 = rowinfo[row0 + r].z, ij1 = rowinfo[row0 + r + 32].z;
            const double v0 = ci.x <= ij0 ? t[r] : 0.0, v1 = ci.x <= ij1 ? t[r + 32] : 0.0;
            p0 = fma(ci.x == ij0 ? 0.5 * v0 : v0, Dab[row0 + r], p0);
            p1 = fma(ci.x == ij1 ? 0.5 * v1 : v1, Dab[row0 + r + 32], p1);
        }
        if (r < ld) {
            const int ij0 = rowinfo[row0 + r].z;
            const double v0 = ci.x <= ij0 ? t[r] : 0.0;
            p0 = fma(ci.x == ij0 ? 0.5 * v0 : v0, Dab[row0 + r], p0);
        }
        double p = p0 + p1;
#pragma unroll
        for (int o = 16; o; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
        if (lane == 0 && p != 0.0) jcol[col] += (ci.y ? 0.5 : 1.0) * p;    // single writer per column; launches are stream-ordered
    }
}

__global__ void jk_gather_kernel(const double *__restrict__ dm, int nao, const int4 *__restrict__ rowinfo, long long nrows, double *Dab,
                                 const int *__restrict__ colc, const int *__restrict__ cold, const int2 *__restrict__ colinfo, long long ncols, double *Dcd)
{
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n < nrows) { const int4 ri = rowinfo[n]; Dab[n] = (ri.w ? 1.0 : 2.0) * dm[(size_t)ri.x * nao + ri.y]; }       // 2 f_IJ D[a,b]
    if (n < ncols) Dcd[n] = (colinfo[n].y ? 1.0 : 2.0) * dm[(size_t)colc[n] * nao + cold[n]];                       // 2 f_KL D[c,d]
}

// scatter the row / column Coulomb parts into J' and symmetrise:  J = J' + J'^T,  K = K' + K'^T
__global__ void jk_scatter_kernel(const int4 *__restrict__ rowinfo, long long nrows, const double *__restrict__ jrow,
                                  const int *__restrict__ colc, const int *__restrict__ cold, long long ncols, const double *__restrict__ jcol,
                                  int nao, double *Jp)
{
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n < nrows && jrow[n] != 0.0) { const int4 ri = rowinfo[n]; atomicAdd(Jp + (size_t)ri.x * nao + ri.y, jrow[n]); }
    if (n < ncols && jcol[n] != 0.0) atomicAdd(Jp + (size_t)colc[n] * nao + cold[n], jcol[n]);
}
__global__ void jk_symm_kernel(const double *__restrict__ P, int nao, double *out)
{
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long long)nao * nao) return;
    const int a = (int)(n / nao), b = (int)(n - (long long)a * nao);
    out[n] = P[n] + P[(size_t)b * nao + a];
}

static inline int shell_dim_sphx(const ShellInfo &s) { return (2 * s.l + 1) * s.nctr; }

static int jk_prepare(CINTOpt *c, JobPlan *plan, DigestState *d)
{
    if (d->jk_ready) return 0;
    if (plan->cart) return b200_fail(CINTB200_ENOSUP, "J/K digestion works on the spherical job");
    const int nbas = c->nbas, nao = c->nao_sph;
    d->nao = nao;
    std::vector<int4> rowinfo(d->nrows);
    for (long long r = 0; r < d->nrows; r++) {
        int i, j;
        pair_shells(plan->row_pair[r], &i, &j);
        const int di = shell_dim_sphx(c->shells[i]);
        rowinfo[r] = make_int4(c->shells[i].ao_sph + plan->row_pos[r] % di, c->shells[j].ao_sph + plan->row_pos[r] / di, plan->row_pair[r], i == j);
    }
    std::vector<int2> colinfo(d->ncols);
    std::vector<int> colc(d->ncols), cold(d->ncols);
    for (long long q = 0; q < d->ncols; q++) {
        int k, l;
        pair_shells(plan->col_pair[q], &k, &l);
        const int dk = shell_dim_sphx(c->shells[k]);
        colinfo[q] = make_int2(plan->col_pair[q], k == l);
        colc[q] = c->shells[k].ao_sph + plan->col_pos[q] % dk;
        cold[q] = c->shells[l].ao_sph + plan->col_pos[q] / dk;
    }
    // outer units and their visit lists (this rank's kets, ascending pair index)
    struct HU { int ao0, nx; std::vector<JKEntry> e; };
    std::vector<std::vector<HU>> byshell(nbas);
    for (int X = 0; X < nbas; X++) {
        const int dx = shell_dim_sphx(c->shells[X]);
        if (dx > 255) return b200_fail(CINTB200_ENOSUP, "J/K digestion: shell %d has %d > 255 functions", X, dx);
        for (int x0 = 0; x0 < dx; x0 += JK_NXMAX) byshell[X].push_back(HU{c->shells[X].ao_sph + x0, std::min(JK_NXMAX, dx - x0), {}});
    }
    for (int k = 0; k < nbas; k++)
        for (int l = 0; l <= k; l++) {
            const int q = k * (k + 1) / 2 + l;
            if (plan->colof[q] < 0) continue;
            if (plan->colof[q] > 0x7fffffffLL) return b200_fail(CINTB200_ENOSUP, "J/K digestion: more than 2^31 tile columns");
            const int dk = shell_dim_sphx(c->shells[k]), dl = shell_dim_sphx(c->shells[l]);
            for (size_t s = 0; s < byshell[k].size(); s++) {        // visit from the first shell: x = c (unit stride), y = d
                const int x0 = byshell[k][s].ao0 - c->shells[k].ao_sph;
                byshell[k][s].e.push_back(JKEntry{(int)plan->colof[q] + x0, q, c->shells[l].ao_sph, dl | dk << 8 | 1 << 16});
            }
            if (l != k)
                for (size_t s = 0; s < byshell[l].size(); s++) {    // visit from the second shell: x = d (stride dk), y = c
                    const int x0 = byshell[l][s].ao0 - c->shells[l].ao_sph;
                    byshell[l][s].e.push_back(JKEntry{(int)plan->colof[q] + x0 * dk, q, c->shells[k].ao_sph, dk | dk << 8});
                }
        }
    std::vector<JKUnit> units;
    std::vector<JKEntry> entries;
    for (int nx = 1; nx <= JK_NXMAX; nx++) {
        d->unit_beg[nx] = (int)units.size();
        for (int X = 0; X < nbas; X++)
            for (HU &hu : byshell[X]) {
                if (hu.nx != nx) continue;
                std::stable_sort(hu.e.begin(), hu.e.end(), [](const JKEntry &p, const JKEntry &q) { return p.kl < q.kl; });
                units.push_back(JKUnit{hu.ao0, hu.nx, (int)entries.size(), (int)(entries.size() + hu.e.size())});
                entries.insert(entries.end(), hu.e.begin(), hu.e.end());
            }
    }
    d->unit_beg[JK_NXMAX + 1] = (int)units.size();
    if (up(&d->d_rowinfo, rowinfo) || up(&d->d_colinfo, colinfo) || up(&d->d_colc, colc) || up(&d->d_cold, cold) ||
        up(&d->d_units, units) || up(&d->d_entries, entries)) return CINTB200_ENOMEM;
    const size_t n2 = (size_t)nao * nao;
    if (b200_dmalloc((void **)&d->d_dm, sizeof(double) * n2) != cudaSuccess || b200_dmalloc((void **)&d->d_Jp, sizeof(double) * n2) != cudaSuccess ||
        b200_dmalloc((void **)&d->d_Kp, sizeof(double) * n2) != cudaSuccess ||
        b200_dmalloc((void **)&d->d_Dab, sizeof(double) * std::max<long long>(1, d->nrows)) != cudaSuccess ||
        b200_dmalloc((void **)&d->d_Dcd, sizeof(double) * std::max<long long>(1, d->ncols)) != cudaSuccess ||
        b200_dmalloc((void **)&d->d_jrow, sizeof(double) * std::max<long long>(1, d->nrows)) != cudaSuccess ||
        b200_dmalloc((void **)&d->d_jcol, sizeof(double) * std::max<long long>(1, d->ncols)) != cudaSuccess ||
        b200_big_alloc((void **)&d->d_PA, sizeof(double) * (size_t)nao * std::max<long long>(1, d->ldmax)) ||
        b200_big_alloc((void **)&d->d_PB, sizeof(double) * (size_t)nao * std::max<long long>(1, d->ldmax)))
        return b200_fail(CINTB200_ENOMEM, "J/K digestion: cannot allocate the work arrays (%zu bytes of row partials)", 2 * sizeof(double) * (size_t)nao * d->ldmax);
    d->jk_ready = 1;
    return 0;
}

// ------------------------------------------------------------------ driver hooks
int digest_begin(CINTOpt *c, JobPlan *plan, const DigestJob &job, const double *dm_dev, cudaStream_t st)
{
    if (!job.checksums && !job.jk) return 0;
    if (plan->rect || (job.jk && plan->ncenter != 4))
        return b200_fail(CINTB200_ENOSUP, "tile consumers need a whole-job plan (J/K: 4-centre only)");
    if (!plan->digest) {
        plan->digest = new DigestState();
        DigestState *d = plan->digest;
        d->nrows = (long long)plan->row_pair.size();
        d->ncols = (long long)plan->col_pair.size();
        for (size_t ch = 0; ch < plan->chunks.size(); ch++)
            d->ldmax = std::max(d->ldmax, plan->rows_before[plan->chunks[ch].second] - plan->rows_before[plan->chunks[ch].first]);
    }
    DigestState *d = plan->digest;
    if (job.checksums) {
        if (checksums_prepare(c, plan, d)) return CINTB200_ENOMEM;
        CU_OK(cudaMemsetAsync(d->d_rowsums, 0, sizeof(double) * 3 * std::max<long long>(1, d->nrows), st));
        d->have_rowsums = 0;
    }
    if (job.jk) {
        if (jk_prepare(c, plan, d)) return CINTB200_ENOMEM;
        const size_t n2 = (size_t)d->nao * d->nao;
        CU_OK(cudaMemcpyAsync(d->d_dm, dm_dev, sizeof(double) * n2, cudaMemcpyDeviceToDevice, st));
        CU_OK(cudaMemsetAsync(d->d_Jp, 0, sizeof(double) * n2, st));
        CU_OK(cudaMemsetAsync(d->d_Kp, 0, sizeof(double) * n2, st));
        CU_OK(cudaMemsetAsync(d->d_jrow, 0, sizeof(double) * std::max<long long>(1, d->nrows), st));
        CU_OK(cudaMemsetAsync(d->d_jcol, 0, sizeof(double) * std::max<long long>(1, d->ncols), st));
        const long long n = std::max(d->nrows, d->ncols);
        jk_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d->d_dm, d->nao, d->d_rowinfo, d->nrows, d->d_Dab, d->d_colc, d->d_cold,
                                                                      d->d_colinfo, d->ncols, d->d_Dcd);
        CU_OK(cudaGetLastError());
    }
    return 0;
}

constexpr int jk_rpt(int nx) { return 2; }          // measured: 4 rows per thread for nx <= 3 is not faster (129 vs 124 ms for nx = 3)
template <int NX>
static void launch_rows(const JKArgs &A, dim3 grid, cudaStream_t st) { jk_rows_kernel<NX, jk_rpt(NX)><<<grid, 128, 0, st>>>(A); }

int digest_tile(CINTOpt *c, JobPlan *plan, const DigestJob &job, int chunk, const double *tile, cudaStream_t st)
{
    if (!job.checksums && !job.jk) return 0;
    DigestState *d = plan->digest;
    const long long row0 = plan->rows_before[plan->chunks[chunk].first];
    const long long ld = plan->rows_before[plan->chunks[chunk].second] - row0;
    const long long ncols = plan->chunk_cols[chunk];
    if (ld == 0 || ncols == 0) return 0;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    if (job.checksums) {
        const unsigned gx = (unsigned)((ld + 255) / 256);
        long long gy = std::max<long long>(1, std::min<long long>(ncols, (sms * 16 + gx - 1) / gx));
        const int cpb = (int)((ncols + gy - 1) / gy);
        gy = (ncols + cpb - 1) / cpb;
        tile_rowsum_kernel<<<dim3(gx, (unsigned)gy), 256, 0, st>>>(tile, ld, ncols, row0, d->d_rowI, d->d_colK, d->d_colg, d->d_rowsums, d->nrows, cpb);
        CU_OK(cudaGetLastError());
        d->have_rowsums = 1;
    }
    if (job.jk) {
        JKArgs A;
        A.tile = tile; A.ld = ld; A.row0 = row0; A.nao = d->nao; A.rowinfo = d->d_rowinfo; A.units = d->d_units; A.entries = d->d_entries;
        A.dm = d->d_dm; A.Dcd = d->d_Dcd; A.PA = d->d_PA; A.PB = d->d_PB; A.jrow = d->d_jrow; A.ldP = d->ldmax; A.want_k = job.want_k;
        for (int nx = 1; nx <= JK_NXMAX; nx++) {
            A.ubeg = d->unit_beg[nx]; A.uend = d->unit_beg[nx + 1];
            if (A.uend <= A.ubeg) continue;
            const unsigned gx = (unsigned)((ld + 128 * jk_rpt(nx) - 1) / (128 * jk_rpt(nx)));
            const unsigned gy = (unsigned)std::max(1, std::min(A.uend - A.ubeg, (sms * 12 + (int)gx - 1) / (int)gx));
            const dim3 grid(gx, gy);
            switch (nx) {
            case 1: launch_rows<1>(A, grid, st); break;
            case 2: launch_rows<2>(A, grid, st); break;
            case 3: launch_rows<3>(A, grid, st); break;
            case 4: launch_rows<4>(A, grid, st); break;
            default: launch_rows<5>(A, grid, st); break;
            }
            CU_OK(cudaGetLastError());
        }
        if (job.want_k) {
            jk_fold_kernel<<<d->nao, 256, sizeof(double) * d->nao, st>>>(d->d_PA, d->d_PB, d->ldmax, ld, row0, d->d_rowinfo, d->nao, d->d_Kp);
            CU_OK(cudaGetLastError());
        }
        jk_cols_kernel<<<sms * 8, 256, 0, st>>>(tile, ld, ncols, row0, d->d_rowinfo, d->d_colinfo, d->d_Dab, d->d_jcol);
        CU_OK(cudaGetLastError());
    }
    return 0;
}

int digest_end(CINTOpt *c, JobPlan *plan, const DigestJob &job, double *vj_dev, double *vk_dev, cudaStream_t st)
{
    (void)c;
    if (!job.jk) return 0;
    DigestState *d = plan->digest;
    const long long n = std::max(d->nrows, d->ncols);
    jk_scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d->d_rowinfo, d->nrows, d->d_jrow, d->d_colc, d->d_cold, d->ncols, d->d_jcol, d->nao, d->d_Jp);
    const long long n2 = (long long)d->nao * d->nao;
    if (vj_dev) jk_symm_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(d->d_Jp, d->nao, vj_dev);
    if (vk_dev && job.want_k) jk_symm_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(d->d_Kp, d->nao, vk_dev);
    CU_OK(cudaGetLastError());
    return 0;
}

void digest_mark_rowsums(JobPlan *plan) { if (plan->digest) plan->digest->have_rowsums = 1; }

// Per-bra-pair fingerprints of the last run with checksums on (this rank's partial sums over its kets):
//   S[p] = sum v, A[p] = sum |v|, F[p] = sum v h(row position) g(c,d)      -- definitions in oracle/ref_golden.c
int digest_fetch_checksums(CINTOpt *c, JobPlan *plan, double *S, double *A, double *F, double *total)
{
    DigestState *d = plan ? plan->digest : nullptr;
    if (!d || !d->have_rowsums) return b200_fail(CINTB200_EINVAL, "no checksums: enable them with cintb200_set_checksums before the whole-job run");
    std::vector<double> rs(3 * (size_t)d->nrows);
    CU_OK(cudaSetDevice(c->device));
    CU_OK(cudaMemcpy(rs.data(), d->d_rowsums, sizeof(double) * rs.size(), cudaMemcpyDeviceToHost));
    const int nb = plan->ncenter == 3 ? plan->aux0 : c->nbas;
    const size_t npair = (size_t)nb * (nb + 1) / 2;
    if (S) memset(S, 0, sizeof(double) * npair);
    if (A) memset(A, 0, sizeof(double) * npair);
    if (F) memset(F, 0, sizeof(double) * npair);
    double tot = 0;
    for (long long r = 0; r < d->nrows; r++) {
        const int p = plan->row_pair[r];
        tot += rs[r];
        if (S) S[p] += rs[r];
        if (A) A[p] += rs[d->nrows + r];
        if (F) F[p] += h_weight(plan->row_pos[r]) * rs[2 * d->nrows + r];
    }
    if (total) *total = tot;
    return (int)npair;
}
