// "Wide" ERI kernel for the high-angular-momentum classes that have no register / cooperative instantiation ((fd|fd), (ff|ff),
// every class with g or h shells): the [e0|f0] block of such a class has thousands of components (63 001 for (hh|hh)), far more
// than a thread or a lane group can hold.
//
// One side of the quartet (the "e side": a pair type (L1 L2), compile time) keeps its components in REGISTERS, in NPARTS slices;
// the other side (the "f side", run time) is spread over threads: a thread owns one unit = (f component, e slice) and, per
// primitive quartet and root, loads the three G columns of its f from shared memory (3 (L1+L2+1) loads) and runs the unrolled
// sum over its e slice (2 FMAs per component) -- the data reuse that the catch-all kernel (kern_generic.cu: one thread per
// component, three shared-memory loads per FMA pair) lacks.  A block handles one (quartet, window of 128 units); windows of a
// quartet recompute roots and the 2-D recurrence (cheap next to the quadrature sum of such classes).  Primitive quartets are
// processed in batches so that the root (2N lanes) and recurrence (3N lanes) phases fill the block.
//
// The block writes its [e0|f0] accumulators to global scratch; HRR, cart->sph and the store are the catch-all kernel's
// epilogue, run in "epilogue only" mode on that scratch (kern_generic.cu).
//
// Reference stages: CINT2e_loop (src/cint2e.c:660-758, incl. the quartet-level test cce_ij + cce_kl > expcutoff :720),
// CINTrys_roots (src/rys_roots.c:57), CINTg0_2e + CINTg0_2e_2d (src/g2e.c:4425, :272), CINTgout2e (src/cint2e.c:961).
// Limits: one contraction combination per pair (general contractions stay with the catch-all kernel), plain and long-range
// Coulomb (the short-range 2N-point rule stays there too).
#include <algorithm>
#include <cstring>
#include "kern_reg.cuh"
#include "kernels.h"
#include "tile_task.cuh"

#define WIDE_THREADS 128

struct WideArgs {
    long long task_base, ntasks;        // this launch handles tasks [task_base, task_base + gridDim.x)
    int la, lb, lc, ld;                 // class in canonical (bra | ket) orientation
    int nroots, nmax, mmax, gs;         // gs = padded (nmax+1)(mmax+1)
    int nE, nF;                         // components of the bra / ket ranges (scratch layout acc[e * nF + f])
    int nf_side, units;                 // components of the run-time side, units per quartet = nf_side * NPARTS
    int pb;                             // primitive quartets per batch
    double *scratch; size_t scratch_per_block;
    int *nonzero;
};

template <int L1, int L2, int NPARTS, bool E_IS_BRA>
__global__ void __launch_bounds__(WIDE_THREADS, 2) eri_wide_kernel(const EngineParams P, const WideArgs W, const Task *__restrict__ tasks, const TileParams TP)
{
    constexpr int NMAXE = L1 + L2;                          // highest degree on the register side
    constexpr int NE = cx_nrange(L1, L1 + L2);
    constexpr int EP = (NE + NPARTS - 1) / NPARTS;          // components per slice
    extern __shared__ double sm[];
    const int tid = threadIdx.x;
    const int N = W.nroots, nmax = W.nmax, mmax = W.mmax, GS = W.gs, PB = W.pb;
    double *s_prep = sm;                                    // [PB][16]
    double *s_rw = s_prep + 16 * PB;                        // [PB][2N]
    double *s_g = s_rw + 2 * N * PB;                        // [PB][3][N][GS]
    const long long t = W.task_base + blockIdx.x;
    if (t >= W.ntasks) return;
    const Task task = tasks ? tasks[t] : tile_task(TP, t);
    double *acc_out = W.scratch + (size_t)blockIdx.x * W.scratch_per_block;
    if (task.bra < 0) return;                               // outside the loop (tile mode): the epilogue skips it too
    const PairHdr hb = P.pairs[task.bra];
    const PairHdr hk = P.pairs[task.ket];
    // my unit
    const int u = blockIdx.y * WIDE_THREADS + tid;
    const bool have = u < W.units;
    const int part = have ? u / W.nf_side : 0, f = have ? u - part * W.nf_side : 0;
    int fx, fy, fz;
    {
        int lf = E_IS_BRA ? W.lc : W.la, r = f;
        while (r >= cx_ncart(lf)) { r -= cx_ncart(lf); lf++; }
        int lx = lf, cnt = 0;
        while (r >= cnt + (lf - lx + 1)) { cnt += lf - lx + 1; lx--; }
        fx = lx; fy = (lf - lx) - (r - cnt); fz = r - cnt;
    }
    // column of (gx, gy, gz) that belongs to my f: the register side runs over n (bra degrees) when it is the bra, else over m
    const int ms = mmax + 1;
    const int ox = E_IS_BRA ? fx : fx * ms, oy = E_IS_BRA ? fy : fy * ms, oz = E_IS_BRA ? fz : fz * ms;
    const int cstride = E_IS_BRA ? ms : 1;
    double acc[EP];
#pragma unroll
    for (int k = 0; k < EP; k++) acc[k] = 0.0;

    const double fsp[2] = {0.282094791773878143, 0.488602511902919921};
    const double common = 34.986836655249725693
        * (W.la < 2 ? fsp[W.la] : 1.0) * (W.lb < 2 ? fsp[W.lb] : 1.0) * (W.lc < 2 ? fsp[W.lc] : 1.0) * (W.ld < 2 ? fsp[W.ld] : 1.0);
    const int npq = hb.npp * hk.npp;
    int executed = 0;
    for (int base = 0; base < npq; base += PB) {
        const int nb = min(PB, npq - base);
        // --- per primitive quartet of the batch: Gaussian products, x, prefactor (one lane each) ---
        if (tid < nb) {
            const int pq = base + tid, kq = pq / hb.npp, bq = pq - kq * hb.npp;
            const PrimPair pb = P.prims[hb.pp_off + bq], pk = P.prims[hk.pp_off + kq];
            const bool ok = pb.cce + pk.cce <= P.expcutoff;            // src/cint2e.c:720
            const double aij = pb.aij, akl = pk.aij, asum = aij + akl, a1 = aij * akl, a0 = a1 / asum;
            const double dx = pb.px - pk.px, dy = pb.py - pk.py, dz = pb.pz - pk.pz;
            double x = a0 * (dx * dx + dy * dy + dz * dz);
            double fac1 = ok ? common * pb.kij * pk.kij * sqrt(a0 / (a1 * a1 * a1)) * P.pcoef[hb.cc_off + bq] * P.pcoef[hk.cc_off + kq] : 0.0;
            double theta = 1.0;
            if (P.omega > 0) { theta = P.omega * P.omega / (P.omega * P.omega + a0); x *= theta; fac1 *= sqrt(theta); }   // src/g2e.c:4477-4492
            double *d = s_prep + 16 * tid;
            d[0] = aij; d[1] = akl; d[2] = x; d[3] = fac1; d[4] = dx; d[5] = dy; d[6] = dz;
            d[7] = pb.px - hb.ra[0]; d[8] = pb.py - hb.ra[1]; d[9] = pb.pz - hb.ra[2];
            d[10] = pk.px - hk.ra[0]; d[11] = pk.py - hk.ra[1]; d[12] = pk.pz - hk.ra[2];
            d[13] = theta; d[14] = ok ? 1.0 : 0.0;
        }
        __syncthreads();
        // --- roots and weights: one polynomial per lane (CINTrys_roots) ---
        for (int w = tid; w < nb * 2 * N; w += WIDE_THREADS) {
            const int p = w / (2 * N), q = w - p * 2 * N;
            s_rw[w] = rys_value(P.rys_coef, N, s_prep[16 * p + 2], q);
        }
        __syncthreads();
        // --- 2-D recurrence: one (primitive, root, axis) per lane (CINTg0_2e_2d) ---
        for (int w = tid; w < nb * 3 * N; w += WIDE_THREADS) {
            const int p = w / (3 * N), rx = w - p * 3 * N, r = rx / 3, xyz = rx - 3 * r;
            const double *d = s_prep + 16 * p;
            const double aij = d[0], akl = d[1], asum = aij + akl;
            const double s = s_rw[p * 2 * N + 2 * r] * d[13];           // t^2 (long range: theta t^2)
            const double sa = s * akl / asum, sk = s * aij / asum;
            const double b00 = 0.5 * s / asum, b10 = 0.5 * (1.0 - sa) / aij, b01 = 0.5 * (1.0 - sk) / akl;
            const double c00 = d[7 + xyz] - sa * d[4 + xyz], c0p = d[10 + xyz] + sk * d[4 + xyz];
            double *g = s_g + (size_t)((p * 3 + xyz) * N + r) * GS;
            g[0] = (xyz == 2) ? s_rw[p * 2 * N + 2 * r + 1] * d[3] : 1.0;
            if (nmax > 0) g[ms] = c00 * g[0];
            for (int n = 1; n < nmax; n++) g[(n + 1) * ms] = c00 * g[n * ms] + n * b10 * g[(n - 1) * ms];
            for (int m = 0; m < mmax; m++)
                for (int n = 0; n <= nmax; n++) {
                    double v = c0p * g[n * ms + m];
                    if (m > 0) v += m * b01 * g[n * ms + m - 1];
                    if (n > 0) v += n * b00 * g[(n - 1) * ms + m];
                    g[n * ms + m + 1] = v;
                }
        }
        __syncthreads();
        // --- quadrature sum over my e slice (CINTgout2e) ---
        if (have) {
            for (int p = 0; p < nb; p++) {
                for (int r = 0; r < N; r++) {
                    const double *gx = s_g + (size_t)((p * 3 + 0) * N + r) * GS + ox;
                    const double *gy = s_g + (size_t)((p * 3 + 1) * N + r) * GS + oy;
                    const double *gz = s_g + (size_t)((p * 3 + 2) * N + r) * GS + oz;
                    double cx[NMAXE + 1], cy[NMAXE + 1], cz[NMAXE + 1];
#pragma unroll
                    for (int n = 0; n <= NMAXE; n++) { cx[n] = gx[n * cstride]; cy[n] = gy[n * cstride]; cz[n] = gz[n * cstride]; }
                    static_for<NPARTS>([&](auto PP) {
                        constexpr int pp = decltype(PP)::value;
                        if (part == pp) {
                            static_for<EP>([&](auto KK) {
                                constexpr int k = decltype(KK)::value, e = pp * EP + k;
                                if constexpr (e < NE) {
                                    constexpr int le = cx_range_l(L1, e), ie = cx_range_i(L1, e);
                                    constexpr int ex = cx_lx(le, ie), ey = cx_ly(le, ie), ez = cx_lz(le, ie);
                                    acc[k] = fma(cx[ex] * cy[ey], cz[ez], acc[k]);
                                }
                            });
                        }
                    });
                }
            }
        }
        if (tid == 0) for (int p = 0; p < nb; p++) executed += s_prep[16 * p + 14] != 0.0;
        __syncthreads();                    // the next batch overwrites the G arrays
    }
    // --- accumulators -> scratch in the catch-all kernel's [e][f] layout (e = bra range index, f = ket range index) ---
    if (have) {
#pragma unroll
        for (int k = 0; k < EP; k++) {
            const int e = part * EP + k;
            if (e < NE) acc_out[E_IS_BRA ? (size_t)e * W.nF + f : (size_t)f * W.nF + e] = acc[k];
        }
    }
    if (tid == 0 && blockIdx.y == 0 && W.nonzero) W.nonzero[t] = executed > 0;
}

// ------------------------------------------------------------------ host side
typedef void (*WideFn)(const EngineParams, const WideArgs, const Task *, const TileParams);
struct WideEntry { int l1, l2, nparts, e_is_bra; WideFn fn; };
#define WIDE_ROW(l1, l2, np) {l1, l2, np, 1, eri_wide_kernel<l1, l2, np, true>}, {l1, l2, np, 0, eri_wide_kernel<l1, l2, np, false>}
// slices keep <= ~80 accumulators per thread: NE(l1 l2) = components of degrees l1 .. l1+l2
static const WideEntry g_wide[] = {
    WIDE_ROW(2, 0, 1), WIDE_ROW(2, 1, 1), WIDE_ROW(2, 2, 1),
    WIDE_ROW(3, 0, 1), WIDE_ROW(3, 1, 1), WIDE_ROW(3, 2, 1), WIDE_ROW(3, 3, 1),
    WIDE_ROW(4, 0, 1), WIDE_ROW(4, 1, 1), WIDE_ROW(4, 2, 1), WIDE_ROW(4, 3, 2), WIDE_ROW(4, 4, 2),
    WIDE_ROW(5, 0, 1), WIDE_ROW(5, 1, 1), WIDE_ROW(5, 2, 2), WIDE_ROW(5, 3, 2), WIDE_ROW(5, 4, 4), WIDE_ROW(5, 5, 4),
};

static int nrange_h(int l0, int l1) { int s = 0; for (int l = l0; l <= l1; l++) s += (l + 1) * (l + 2) / 2; return s; }

// Choose the orientation for class (la lb | lc ld); returns the table entry or NULL when the class is better left to the
// catch-all kernel (small blocks, no instantiation).  units = run-time-side components x slices.
static const WideEntry *wide_choose(int la, int lb, int lc, int ld, int *units, int *nf_side)
{
    const WideEntry *best = nullptr;
    double best_score = 0;
    for (const WideEntry &e : g_wide) {
        const int l1 = e.e_is_bra ? la : lc, l2 = e.e_is_bra ? lb : ld, f1 = e.e_is_bra ? lc : la, f2 = e.e_is_bra ? ld : lb;
        if (e.l1 != l1 || e.l2 != l2) continue;
        const int ne = nrange_h(l1, l1 + l2), nf = nrange_h(f1, f1 + f2), u = nf * e.nparts;
        if (u < 40) continue;                                   // too few units to fill a block
        const int ep = (ne + e.nparts - 1) / e.nparts;
        // useful FMA pairs per shared-memory load, discounted by idle threads of the last window
        const int windows = (u + WIDE_THREADS - 1) / WIDE_THREADS;
        const double score = (double)ep / (3.0 * (l1 + l2 + 1)) * u / (windows * (double)WIDE_THREADS);
        if (score > best_score) { best_score = score; best = &e; *units = u; *nf_side = nf; }
    }
    return best;
}

int wide_eligible(int la, int lb, int lc, int ld, int ncab, int nccd, int short_range)
{
    if (ncab * nccd != 1 || short_range) return 0;
    const long long nef = (long long)nrange_h(la, la + lb) * nrange_h(lc, lc + ld);
    if (nef < 600) return 0;                                    // small blocks: the catch-all kernel is fine
    int u, nf;
    return wide_choose(la, lb, lc, ld, &u, &nf) != nullptr;
}

// Quadrature part of tasks [task_base, task_base + ntask_here) into scratch (one scratch block per task, stride scratch_per_block)
int wide_launch(const EngineParams &P, const GenericClass &C, const Task *tasks, long long task_base, long long ntasks, int ntask_here,
                int *nonzero, cudaStream_t stream, const TileParams *tile)
{
    int units = 0, nf_side = 0;
    const WideEntry *e = wide_choose(C.la, C.lb, C.lc, C.ld, &units, &nf_side);
    if (!e) return -1;
    WideArgs W;
    memset(&W, 0, sizeof W);
    W.task_base = task_base; W.ntasks = ntasks;
    W.la = C.la; W.lb = C.lb; W.lc = C.lc; W.ld = C.ld;
    W.nroots = C.nroots; W.nmax = C.la + C.lb; W.mmax = C.lc + C.ld;
    W.gs = ((W.nmax + 1) * (W.mmax + 1)) | 1;
    W.nE = C.nE; W.nF = C.nF; W.nf_side = nf_side; W.units = units;
    W.pb = std::max(1, std::min(8, WIDE_THREADS / (3 * C.nroots)));
    size_t smem;
    for (;;) {
        smem = sizeof(double) * ((size_t)16 * W.pb + (size_t)2 * C.nroots * W.pb + (size_t)W.pb * 3 * C.nroots * W.gs);
        if (smem <= 100 * 1024 || W.pb == 1) break;             // two blocks per SM
        W.pb--;
    }
    W.scratch = C.scratch; W.scratch_per_block = C.scratch_per_block; W.nonzero = nonzero;
    TileParams TP;
    memset(&TP, 0, sizeof TP);
    if (tile) TP = *tile;
    if (smem > 48 * 1024 && cudaFuncSetAttribute((const void *)e->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    const dim3 grid((unsigned)ntask_here, (unsigned)((units + WIDE_THREADS - 1) / WIDE_THREADS));
    void *args[] = {(void *)&P, (void *)&W, (void *)&tasks, (void *)&TP};
    return cudaLaunchKernel((const void *)e->fn, grid, dim3(WIDE_THREADS), args, smem, stream) == cudaSuccess ? 0 : -1;
}
