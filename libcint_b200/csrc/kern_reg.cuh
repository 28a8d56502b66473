// Register-resident ERI kernel: ONE THREAD PER SHELL QUARTET, fully specialised at compile time on
// the angular class (LA LB | LC LD) and on the number of contraction combinations of each pair.
//
// Used by the whole-job / tile driver (driver.cu) for every class whose [e0|f0] block fits the
// register file.  Geometry of a launch ("tile"):
//   T side = bra pairs (ab|, one per THREAD, consecutive threads = consecutive pairs of one pair
//            class, primitive data read from a structure-of-arrays table -> fully coalesced loads;
//   U side = ket pair |cd), UNIFORM per block (blockIdx.y), primitive data staged once in smem and
//            read back as broadcasts.
// so the primitive loops have block-uniform trip counts (T pairs are padded to the class maximum
// with zero-weight primitives) and there is no divergence.  Output goes to a column-major tile
// out[row(ab) + ld * col(cd)] -- consecutive threads write consecutive rows.
//
// Reference stages replaced (same math, different order of operations):
//   CINT2e_loop (src/cint2e.c:660-758)   -> the two primitive loops below.  Pair-level screening only: the quartet-level
//                                           test cce_ij + cce_kl > expcutoff (:720) would drop 3.6% more primitives, all
//                                           < e^-60.  A warp-uniform form of it (primitives sorted by cce, surviving U
//                                           primitives = a prefix bounded by the warp's smallest cce_ij) was built and
//                                           measured in round 2: C60 1137 -> 1190 ms, the bound costs more than it saves
//                                           (uncontracted and cooperative classes +4-18%, the contracted ones -1%)
//   CINTrys_roots (src/rys_roots.c:57)   -> rys_roots_t2w<N> from the smem-staged table
//   CINTg0_2e (src/g2e.c:4518-4540)      -> b00/b10/b01/c00/c0p written in t^2, no division
//   CINTg0_2e_2d (src/g2e.c:272-421)     -> register VRR, unrolled
//   CINTg0_*2d_4d + CINTgout2e           -> [e0|f0] accumulation per root, HRR once per contracted quartet
//   CINTprim_to_ctr_0/1 (src/g1e.c:530)  -> coefficient products from the pair table
//   c2s_sph_2e1 (src/cart2sph.c:5324)    -> constexpr sparse transforms + direct strided store
#pragma once
#include <utility>
#include <type_traits>
#include "types.h"
#include "rys.cuh"
#include "c2s_constexpr.inc"
#include "tune.inc"

// ----------------------------------------------------------------------------- compile-time helpers
template <class F, int... I>
__device__ __forceinline__ void static_for_impl(F &&f, std::integer_sequence<int, I...>)
{
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F &&f)
{
    static_for_impl(static_cast<F &&>(f), std::make_integer_sequence<int, N>{});
}

__host__ __device__ constexpr int cx_ncart(int l) { return (l + 1) * (l + 2) / 2; }
__host__ __device__ constexpr int cx_lx(int l, int idx)
{
    int n = 0;
    for (int lx = l; lx >= 0; lx--)
        for (int ly = l - lx; ly >= 0; ly--, n++)
            if (n == idx) return lx;
    return -1;
}
__host__ __device__ constexpr int cx_ly(int l, int idx)
{
    int n = 0;
    for (int lx = l; lx >= 0; lx--)
        for (int ly = l - lx; ly >= 0; ly--, n++)
            if (n == idx) return ly;
    return -1;
}
__host__ __device__ constexpr int cx_lz(int l, int idx) { return l - cx_lx(l, idx) - cx_ly(l, idx); }
__host__ __device__ constexpr int cx_idx(int l, int lx, int lz) { return (l - lx) * (l - lx + 1) / 2 + lz; }
// number of components of degrees l0..l1
__host__ __device__ constexpr int cx_nrange(int l0, int l1)
{
    int s = 0;
    for (int l = l0; l <= l1; l++) s += cx_ncart(l);
    return s;
}
// degree / sub-index of component e of the range starting at l0
__host__ __device__ constexpr int cx_range_l(int l0, int e)
{
    int l = l0;
    while (e >= cx_ncart(l)) { e -= cx_ncart(l); l++; }
    return l;
}
__host__ __device__ constexpr int cx_range_i(int l0, int e)
{
    int l = l0;
    while (e >= cx_ncart(l)) { e -= cx_ncart(l); l++; }
    return e;
}
// size of HRR level J: sum_{le=L0}^{L0+LB-J} ncart(le)*ncart(J)
__host__ __device__ constexpr int cx_hrr_size(int l0, int lb, int j)
{
    int s = 0;
    for (int le = l0; le <= l0 + lb - j; le++) s += cx_ncart(le) * cx_ncart(j);
    return s;
}
__host__ __device__ constexpr int cx_hrr_off(int l0, int le, int j)   // offset of block le inside level j
{
    int s = 0;
    for (int l = l0; l < le; l++) s += cx_ncart(l) * cx_ncart(j);
    return s;
}


#define REG_THREADS 128
#ifndef REG_ACC_SMEM_MAX
#define REG_ACC_SMEM_MAX 0       // >0: keep outer accumulators of contracted-T classes in smem (measured: no gain, off)
#endif
#ifndef REG_UQ_UNROLL
#define REG_UQ_UNROLL 1          // unroll factor of the inner (U) primitive loop
#endif
#ifndef REG_MIN_BLOCKS
#define REG_MIN_BLOCKS 2          // register cap 255; see profiles/ for the occupancy experiments
#endif
#define REG_MAXU 64            // most primitive pairs of a U pair (8 x 8)

// doubles of shared memory per warp for the epilogue's output staging: [32][RB] values + [32][RB] int2 table + [32] int2
__host__ __device__ constexpr int reg_stage_doubles(int rb) { return 32 * rb * 2 + 32; }

// row stride (doubles) of the smem copy of the Rys table: odd, so rows fall on distinct 8-byte bank pairs
__host__ __device__ constexpr int rys_smem_stride(int n) { return (RYS_DEG + 1) * 2 * n + 1; }
// the register kernels of order <= RYS_FNMAX use the degree-6 table set (rys.cuh)
__host__ __device__ constexpr bool reg_fast_rys(int n) { return REG_FAST_RYS && n <= REG_FAST_NMAX && n <= RYS_FNMAX; }
__host__ __device__ constexpr int reg_rys_row(int n) { return (reg_fast_rys(n) ? RYS_FDEG + 1 : RYS_DEG + 1) * 2 * n; }
// REG_RYS_VEC2: the t^2 and w coefficients of one root and degree are neighbours in a table row, so one 16-byte load fetches
// both (half the LDS instructions, ~20% fewer shared-memory wavefronts for random rows).  Rows then need 16-byte alignment
// and an ODD stride in 16-byte units (conflict-free quarter-warps) instead of an odd stride in doubles.
#ifndef REG_RYS_VEC2
#define REG_RYS_VEC2 0
#endif
__host__ __device__ constexpr int reg_rys_stride(int n)
{
    return REG_RYS_VEC2 ? (((reg_rys_row(n) / 2) & 1) ? reg_rys_row(n) : reg_rys_row(n) + 2) : reg_rys_row(n) + 1;
}

template <int N>
__device__ __forceinline__ void rys_roots_smem_fast(const double *tab, double x, double (&t2)[N], double (&w)[N])
{
    // same branch-free structure as rys_roots_smem below, degree 6 on the 64-intervals-per-octave grid
    const bool large = x >= 35.0 + 5.0 * N;
    const double isx = fast_rsqrt(large ? x : 1.0);
    const double ix = isx * isx;
    int idx;
    double y;
    rys_locate_m<RYS_FM>(large ? 0.0 : x, idx, y);
    // keep the interval index opaque: otherwise the grid's exponent offset (bits(8.0) >> 46, times the row stride) is folded
    // into the LDS immediates, overflows their 24 bits and is re-materialised with one integer add PER coefficient load
    asm("" : "+r"(idx));
    const double *c = tab + idx * reg_rys_stride(N);
    static_assert(RYS_FDEG == 6, "Estrin scheme below is written for degree 6");
    const double y2 = y * y, y4 = y2 * y2;
#if REG_RYS_VEC2
    const double2 *c2 = reinterpret_cast<const double2 *>(c);          // [j][k] = {t^2 coefficient, w coefficient}
#pragma unroll
    for (int k = 0; k < N; k++) {
        const double2 a0 = c2[0 * N + k], a1 = c2[1 * N + k], a2 = c2[2 * N + k], a3 = c2[3 * N + k];
        const double2 a4 = c2[4 * N + k], a5 = c2[5 * N + k], a6 = c2[6 * N + k];
        const double tv = fma(fma(a6.x, y2, fma(a5.x, y, a4.x)), y4, fma(fma(a3.x, y, a2.x), y2, fma(a1.x, y, a0.x)));
        const double wv = fma(fma(a6.y, y2, fma(a5.y, y, a4.y)), y4, fma(fma(a3.y, y, a2.y), y2, fma(a1.y, y, a0.y)));
        t2[k] = large ? c_rys_lx_r[N * (N - 1) / 2 + k] * ix : tv;
        w[k] = large ? c_rys_lx_v[N * (N - 1) / 2 + k] * isx : wv;
    }
    return;
#endif
#pragma unroll
    for (int p = 0; p < 2 * N; p++) {
        const double p01 = fma(c[1 * 2 * N + p], y, c[0 * 2 * N + p]);
        const double p23 = fma(c[3 * 2 * N + p], y, c[2 * 2 * N + p]);
        const double p45 = fma(c[5 * 2 * N + p], y, c[4 * 2 * N + p]);
        const double q0 = fma(p23, y2, p01), q1 = fma(c[6 * 2 * N + p], y2, p45);
        const double v = fma(q1, y4, q0);
        if (p & 1) w[p >> 1] = large ? c_rys_lx_v[N * (N - 1) / 2 + (p >> 1)] * isx : v;
        else t2[p >> 1] = large ? c_rys_lx_r[N * (N - 1) / 2 + (p >> 1)] * ix : v;
    }
}

template <int N>
__device__ __forceinline__ void rys_roots_smem(const double *tab, int nint, double x, double (&t2)[N], double (&w)[N])
{
    // BRANCH-FREE on purpose: the table polynomials and the large-x form are both evaluated and selected, so the
    // primitive loop body stays one basic block and the scheduler can overlap this latency-bound chain with the
    // VRR / quadrature arithmetic of the neighbouring iteration (ncu: these kernels stall on fixed-latency
    // dependencies, not on FP64 issue bandwidth).  Cost: one rsqrt and 2N multiplies per primitive quartet.
    const bool large = x >= 35.0 + 5.0 * N;
    const double isx = fast_rsqrt(large ? x : 1.0);          // t^2 = r/x, w = v/sqrt(x)
    const double ix = isx * isx;
    int idx;
    double y;
    rys_locate(large ? 0.0 : x, idx, y);
    (void)nint;
    asm("" : "+r"(idx));            // see rys_roots_smem_fast: keeps the LDS immediates small (matters from nroots = 4 on)
    const double *c = tab + idx * reg_rys_stride(N);
    static_assert(RYS_DEG == 9, "Estrin scheme below is written for degree 9");
    // Estrin evaluation: depth 4 instead of Horner's 9 dependent FMAs
    const double y2 = y * y, y4 = y2 * y2, y8 = y4 * y4;
#if REG_RYS_VEC2
    const double2 *c2 = reinterpret_cast<const double2 *>(c);
#pragma unroll
    for (int k = 0; k < N; k++) {
        const double2 a0 = c2[0 * N + k], a1 = c2[1 * N + k], a2 = c2[2 * N + k], a3 = c2[3 * N + k], a4 = c2[4 * N + k];
        const double2 a5 = c2[5 * N + k], a6 = c2[6 * N + k], a7 = c2[7 * N + k], a8 = c2[8 * N + k], a9 = c2[9 * N + k];
        const double tv = fma(fma(a9.x, y, a8.x), y8, fma(fma(fma(a7.x, y, a6.x), y2, fma(a5.x, y, a4.x)), y4, fma(fma(a3.x, y, a2.x), y2, fma(a1.x, y, a0.x))));
        const double wv = fma(fma(a9.y, y, a8.y), y8, fma(fma(fma(a7.y, y, a6.y), y2, fma(a5.y, y, a4.y)), y4, fma(fma(a3.y, y, a2.y), y2, fma(a1.y, y, a0.y))));
        t2[k] = large ? c_rys_lx_r[N * (N - 1) / 2 + k] * ix : tv;
        w[k] = large ? c_rys_lx_v[N * (N - 1) / 2 + k] * isx : wv;
    }
    return;
#endif
#pragma unroll
    for (int p = 0; p < 2 * N; p++) {
        const double p01 = fma(c[1 * 2 * N + p], y, c[0 * 2 * N + p]);
        const double p23 = fma(c[3 * 2 * N + p], y, c[2 * 2 * N + p]);
        const double p45 = fma(c[5 * 2 * N + p], y, c[4 * 2 * N + p]);
        const double p67 = fma(c[7 * 2 * N + p], y, c[6 * 2 * N + p]);
        const double p89 = fma(c[9 * 2 * N + p], y, c[8 * 2 * N + p]);
        const double q0 = fma(p23, y2, p01), q1 = fma(p67, y2, p45);
        const double r0 = fma(q1, y4, q0);
        const double v = fma(p89, y8, r0);
        if (p & 1) w[p >> 1] = large ? c_rys_lx_v[N * (N - 1) / 2 + (p >> 1)] * isx : v;
        else t2[p >> 1] = large ? c_rys_lx_r[N * (N - 1) / 2 + (p >> 1)] * ix : v;
    }
}

// ----------------------------------------------------------------------------- HRR in registers
// in : level J-1, layout [PRE][part(J-1)][POST];  out: level J.   (a, b+1_d| = (a+1_d, b| + AB_d (a, b|
template <int L0, int LB, int J, int PRE, int POST>
__device__ __forceinline__ void hrr_step_reg(const double *in, double *out, const double (&ab)[3])
{
    constexpr int NB_IN = cx_ncart(J - 1), NB_OUT = cx_ncart(J);
    constexpr int IN_PART = cx_hrr_size(L0, LB, J - 1), OUT_PART = cx_hrr_size(L0, LB, J);
    static_for<LB - J + 1>([&](auto LEI) {
        constexpr int le = L0 + decltype(LEI)::value;
        static_for<cx_ncart(le)>([&](auto IE) {
            constexpr int ie = decltype(IE)::value;
            static_for<NB_OUT>([&](auto IB) {
                constexpr int ib = decltype(IB)::value;
                constexpr int bx = cx_lx(J, ib), by = cx_ly(J, ib), bz = cx_lz(J, ib);
                constexpr int d = bx ? 0 : (by ? 1 : 2);
                constexpr int ibp = cx_idx(J - 1, bx - (d == 0), bz - (d == 2));
                constexpr int ax = cx_lx(le, ie), az = cx_lz(le, ie);
                constexpr int iep = cx_idx(le + 1, ax + (d == 0), az + (d == 2));
                constexpr int o_out = cx_hrr_off(L0, le, J) + ie * NB_OUT + ib;
                constexpr int o_lo = cx_hrr_off(L0, le, J - 1) + ie * NB_IN + ibp;
                constexpr int o_hi = cx_hrr_off(L0, le + 1, J - 1) + iep * NB_IN + ibp;
                static_for<PRE>([&](auto PP) {
                    constexpr int p = decltype(PP)::value;
                    static_for<POST>([&](auto QQ) {
                        constexpr int q = decltype(QQ)::value;
                        out[(p * OUT_PART + o_out) * POST + q] =
                            fma(ab[d], in[(p * IN_PART + o_lo) * POST + q], in[(p * IN_PART + o_hi) * POST + q]);
                    });
                });
            });
        });
    });
}

// levels J..LB of the HRR, one temporary per intermediate level (any LB)
template <int L0, int LB, int J, int PRE, int POST>
__device__ __forceinline__ void hrr_chain_reg(const double *in, double *out, const double (&ab)[3])
{
    if constexpr (J == LB) {
        hrr_step_reg<L0, LB, J, PRE, POST>(in, out, ab);
    } else {
        double tmp[PRE * cx_hrr_size(L0, LB, J) * POST];
        hrr_step_reg<L0, LB, J, PRE, POST>(in, tmp, ab);
        hrr_chain_reg<L0, LB, J + 1, PRE, POST>(tmp, out, ab);
    }
}

// full HRR of one pair: in [PRE][range L0..L0+LB][POST] -> out [PRE][ncart(L0)][ncart(LB)][POST]
template <int L0, int LB, int PRE, int POST>
__device__ __forceinline__ void hrr_pair_reg(const double *in, double *out, const double (&ab)[3])
{
    if constexpr (LB == 0) {
        static_for<PRE * cx_ncart(L0) * POST>([&](auto I) { out[decltype(I)::value] = in[decltype(I)::value]; });
    } else {
        hrr_chain_reg<L0, LB, 1, PRE, POST>(in, out, ab);
    }
}

// cart -> real spherical on one index of a register array: in [PRE][ncart(L)][POST] -> out [PRE][2L+1][POST]
template <int L, int PRE, int POST>
__device__ __forceinline__ void c2s_reg(const double *in, double *out)
{
    static_assert(L <= C2S_CX_LMAX, "constexpr c2s table too small");
    constexpr int NC = cx_ncart(L), NS = 2 * L + 1;
    static_for<PRE>([&](auto PP) {
        constexpr int p = decltype(PP)::value;
        static_for<NS>([&](auto MM) {
            constexpr int m = decltype(MM)::value;
            static_for<POST>([&](auto QQ) {
                constexpr int q = decltype(QQ)::value;
                double s = 0;
                static_for<NC>([&](auto CC) {
                    constexpr int c = decltype(CC)::value;
                    constexpr double coef = c2s_cx(L, m, c);
                    if constexpr (coef != 0.0) s = fma(coef, in[(p * NC + c) * POST + q], s);
                });
                out[(p * NS + m) * POST + q] = s;
            });
        });
    });
}

template <int L> struct SphDim { static constexpr int value = (L < 2) ? cx_ncart(L) : 2 * L + 1; };

// ----------------------------------------------------------------------------- the kernel
// Classes whose T pair is generally contracted (NCT > 1) update the outer accumulators only once per T primitive:
// they are kept in SHARED memory (layout [i][thread], conflict-free) when small enough, which frees 2*NACC registers
// and lets three blocks share an SM instead of two.
__host__ __device__ constexpr bool reg_acc_in_smem(int nct, int nacc) { return nct > 1 && nacc <= REG_ACC_SMEM_MAX; }
__host__ __device__ constexpr int reg_min_blocks(int la, int lb, int lc, int ld, int nct, int ncu)
{
    return tune_lookup(g_reg_tune, sizeof(g_reg_tune) / sizeof(ClassTune), tune_key(la, lb, lc, ld, nct, ncu), false, REG_MIN_BLOCKS);
}
__host__ __device__ constexpr int reg_uq_unroll(int la, int lb, int lc, int ld, int nct, int ncu)
{
    return tune_lookup(g_reg_tune, sizeof(g_reg_tune) / sizeof(ClassTune), tune_key(la, lb, lc, ld, nct, ncu), true, REG_UQ_UNROLL);
}

template <int LA, int LB, int LC, int LD, int NCT, int NCU, bool RS = false, bool CART = false>
__global__ void __launch_bounds__(REG_THREADS, reg_acc_in_smem(NCT, NCT * NCU * cx_nrange(LA, LA + LB) * cx_nrange(LC, LC + LD)) ? 3 : reg_min_blocks(LA, LB, LC, LD, NCT, NCU))
eri_reg_kernel(const TileParams P)
{
    constexpr int NMAX = LA + LB, MMAX = LC + LD;
    constexpr int N = (LA + LB + LC + LD) / 2 + 1;
    constexpr int NE = cx_nrange(LA, LA + LB), NF = cx_nrange(LC, LC + LD), NEF = NE * NF;
    constexpr int NFA = cx_ncart(LA), NFB = cx_ncart(LB), NFC = cx_ncart(LC), NFD = cx_ncart(LD);
    // CART: Cartesian output (int2e_cart, c2s_cart_2e1 src/cart2sph.c:5845): the cart->sph stages of the epilogue are skipped
    constexpr int DA = CART ? NFA : SphDim<LA>::value, DB = CART ? NFB : SphDim<LB>::value;
    constexpr int DC = CART ? NFC : SphDim<LC>::value, DD = CART ? NFD : SphDim<LD>::value;
    constexpr int USTR = 9 + NCU;      // doubles per staged U primitive

    extern __shared__ double smem[];
    double *s_rys = smem;                                       // [nint][stride]
    const int tid = threadIdx.x;

    // --- stage the Rys table of N roots ONCE per block; the block then walks a contiguous range of work items
    //     (item = one ket x 32 bras, one WARP each), so the table and the ket's primitives are reused ---
    constexpr bool FASTRYS = reg_fast_rys(N);
    constexpr int RSTR = reg_rys_stride(N);
    const int nint = FASTRYS ? c_rys_meta.fast_nint[N] : c_rys_meta.nint[N];
    {
        constexpr int ROW = reg_rys_row(N);
        for (int i = tid; i < nint * ROW; i += REG_THREADS) {
            int r = i / ROW, c = i - r * ROW;
            s_rys[r * RSTR + c] = __ldg(P.rys + i);
        }
    }
    constexpr int NWARP = REG_THREADS / 32;
    const int warp = tid >> 5, lane = tid & 31;
    double *s_u = s_rys + nint * RSTR + (size_t)warp * P.umax * USTR;     // this WARP's ket primitives [nppu <= umax][USTR]
    // per-warp output staging (epilogue): one column of the 32 quartets' blocks [32][RB] + gather table + per-thread row info
    constexpr int RB = DA * DB;
    double *s_st = s_rys + nint * RSTR + (size_t)NWARP * P.umax * USTR
                   + (reg_acc_in_smem(NCT, NCT * NCU * NEF) ? (size_t)NCT * NCU * NEF * REG_THREADS : 0) + (size_t)warp * reg_stage_doubles(RB);
    int2 *s_tab = (int2 *)(s_st + 32 * RB);                     // [32*RB] {row offset in the tile or -1, index into s_st}
    int2 *s_meta = s_tab + 32 * RB;                             // [32] {row base or -1, +di if a is the first index else -di}
    const long long total = P.items ? P.nitems : (long long)P.gx * P.NU;
    int cur_by = -1, t_lo = P.t_begin, t_hi_k = P.t_end;
    PairHdr hu;
    __syncthreads();                    // table staged; from here on the warps run independently (no block barriers)
    // dynamic scheduling per WARP: a warp grabs batches of P.batch consecutive work items (item = one ket x 32 bras)
    // from the per-launch counter; consecutive items share the ket, whose primitives sit in the warp's own smem slice
    for (;;) {
    long long item0 = 0;
    if (lane == 0) item0 = (long long)atomicAdd(P.counter, (unsigned int)P.batch);
    item0 = __shfl_sync(0xffffffffu, item0, 0);
    if (item0 >= total) break;
    const long long item1 = (item0 + P.batch < total) ? item0 + P.batch : total;
    for (long long item = item0; item < item1; item++) {
    int by, bx = 0, u, t_hi = P.t_end, t0l = 0;
    if (P.items) {                      // list mode: explicit items, consecutive items of a ket share its staged primitives
        const int4 it = P.items[item];
        by = u = it.x; t0l = it.y; t_hi = it.y + it.z;
    } else {
        by = (int)(item / P.gx); bx = (int)(item - (long long)by * P.gx);
        u = P.u_first + P.u_step * by;
    }
    if (by != cur_by) {
        __syncwarp();                   // previous ket's primitives no longer in use by this warp
        hu = P.pairs[P.upair[u]];
        for (int i = lane; i < hu.npp; i += 32) {
            const PrimPair pp = P.prims[hu.pp_off + i];
            double *d = s_u + i * USTR;
            d[0] = pp.aij; d[1] = pp.inv_aij; d[2] = pp.px; d[3] = pp.py; d[4] = pp.pz;
            d[5] = pp.px - hu.ra[0]; d[6] = pp.py - hu.ra[1]; d[7] = pp.pz - hu.ra[2];
            if constexpr (NCU == 1) {
                d[8] = pp.kij * P.pcoef[hu.cc_off + i];
                d[9] = 1.0;
            } else {
                d[8] = pp.kij;
                for (int c = 0; c < NCU; c++) d[9 + c] = P.pcoef[hu.cc_off + i * NCU + c];
            }
        }
        __syncwarp();
        cur_by = by;
        // --- which T pairs do this ket's items cover? (once per ket) ---
        // reference loop bound k <= i (examples/time_c60.c:206).  tri = 0: this ket lies below the chunk's bra shells,
        // every T pair is valid (list sorted by primitive count).  tri = 1: the list is sorted by the bra's larger
        // shell index and the valid T pairs are the suffix starting at the first pair with I >= K.
        // tri = 2: one launch serves both kinds of ket -- below the chunk's bra shells (range as given, ordering A) and inside
        // them (suffix of the same range in ordering B, tB rows further down the tables)
        t_lo = P.t_begin; t_hi_k = P.t_end;
        if (P.tri) {
            const int K = P.uK[u];
            if (P.tri == 1 || K >= P.tri_i0) {
                int lo = P.t_begin + P.tB, hi = P.t_end + P.tB;
                t_hi_k = hi;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (P.tI[mid] < K) lo = mid + 1; else hi = mid;
                }
                t_lo = lo;
            }
        }
    }
    if (!P.items) t_hi = t_hi_k;
    const int t0 = P.items ? t0l : t_lo + bx * 32;
    if (t0 >= t_hi) continue;           // warp-uniform
    const int t = t0 + lane;
    const bool active = t < t_hi;
    const int to = active ? t : t_hi - 1;                   // my T occurrence (output placement)
    const int tt = P.tsel ? P.tsel[to] : to;                // its row in the class' pair table
    // warp-uniform primitive loop bound: the largest count among the warp's T pairs (shorter pairs are padded with
    // zero-weight primitives; neighbouring list entries have similar counts by construction)
    // Schwarz: if every quartet of this warp is bounded below the threshold, skip the primitive loops -- the
    // accumulators stay zero and the epilogue zero-fills the blocks (what the reference does for empty blocks)
    const size_t NT = P.NT;
    const bool negligible = P.schwarz_thr > 0 && P.tq[tt] * P.uq[u] < P.schwarz_thr;
    const int Qb = __all_sync(0xffffffffu, negligible) ? 0 : __reduce_max_sync(0xffffffffu, P.tnpp[tt]);

    // --- per-thread (T pair) constants ---
    double ra[3], abT[3];
#pragma unroll
    for (int d = 0; d < 3; d++) { ra[d] = P.tgeom[d * NT + tt]; abT[d] = P.tgeom[(3 + d) * NT + tt]; }
    const double abU[3] = {hu.ab[0], hu.ab[1], hu.ab[2]};
    constexpr double fsp0 = 0.282094791773878143, fsp1 = 0.488602511902919921;
    constexpr double common = 34.986836655249725693
        * (LA == 0 ? fsp0 : LA == 1 ? fsp1 : 1.0) * (LB == 0 ? fsp0 : LB == 1 ? fsp1 : 1.0)
        * (LC == 0 ? fsp0 : LC == 1 ? fsp1 : 1.0) * (LD == 0 ? fsp0 : LD == 1 ? fsp1 : 1.0);

    constexpr int NACC = NCT * NCU * NEF;
    constexpr bool ACC_SMEM = reg_acc_in_smem(NCT, NACC);
    double acc[ACC_SMEM ? 1 : NACC];
    double *s_acc = s_rys + nint * RSTR + (size_t)NWARP * P.umax * USTR + tid;         // [NACC][REG_THREADS]
    if constexpr (ACC_SMEM) {
#pragma unroll
        for (int i = 0; i < NACC; i++) s_acc[i * REG_THREADS] = 0.0;
    } else {
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] = 0.0;
    }

    const int nppu = hu.npp;
    for (int tq = 0; tq < Qb; tq++) {
        const size_t o = (size_t)tq * NT + tt;
        const size_t F = (size_t)P.Q * NT;
        const double aT = P.tprim[o], iaT = P.tprim[F + o];
        const double ptx = P.tprim[2 * F + o], pty = P.tprim[3 * F + o], ptz = P.tprim[4 * F + o];
        double kT = P.tprim[5 * F + o];
        double ccT[NCT];
        if constexpr (NCT == 1) { kT *= P.tprim[6 * F + o]; ccT[0] = 1.0; }
        else {
#pragma unroll
            for (int c = 0; c < NCT; c++) ccT[c] = P.tprim[(6 + c) * F + o];
        }
        const double pa[3] = {ptx - ra[0], pty - ra[1], ptz - ra[2]};
        // accumulator of the inner (U) loop; aliases acc when the T side is uncontracted
        double accu[(NCT == 1) ? 1 : NCU * NEF];
        if constexpr (NCT > 1) {
#pragma unroll
            for (int i = 0; i < NCU * NEF; i++) accu[i] = 0.0;
        }
        constexpr int UQ_UNROLL = reg_uq_unroll(LA, LB, LC, LD, NCT, NCU);
#pragma unroll UQ_UNROLL
        for (int uq = 0; uq < nppu; uq++) {
            const double *su = s_u + uq * USTR;
            const double aU = su[0], iaU = su[1];
            const double asum = aT + aU;
            const double rs = fast_rsqrt(asum);
            const double inv = rs * rs;
            const double pq[3] = {ptx - su[2], pty - su[3], ptz - su[4]};
            const double qc[3] = {su[5], su[6], su[7]};
            const double a0 = aT * aU * inv;
            const double x = a0 * (pq[0] * pq[0] + pq[1] * pq[1] + pq[2] * pq[2]);
            const double fac = common * kT * su[8] * iaT * iaU * rs;
            const double rho_u = aU * inv, rho_t = aT * inv;
            double val[(NCU == 1) ? 1 : NEF];
            if constexpr (NCU > 1) {
#pragma unroll
                for (int i = 0; i < NEF; i++) val[i] = 0.0;
            }
            // one quadrature rule: roots at xx, weights scaled by ff, t^2 scaled by thp (1 unless range-separated)
            auto quad = [&](const double xx, const double ff, const double thp) {
            double t2[N], w[N];
            if constexpr (FASTRYS) rys_roots_smem_fast<N>(s_rys, xx, t2, w);
            else rys_roots_smem<N>(s_rys, nint, xx, t2, w);
            static_for<N>([&](auto RR) {
                constexpr int r = decltype(RR)::value;
                const double s = RS ? t2[r] * thp : t2[r];
                const double su_ = s * rho_u, st_ = s * rho_t;
                const double b00 = 0.5 * s * inv;
                const double b10 = (0.5 - 0.5 * su_) * iaT;
                const double b01 = (0.5 - 0.5 * st_) * iaU;
                double g[3][NMAX + 1][MMAX + 1];
                static_for<3>([&](auto DDm) {
                    constexpr int d = decltype(DDm)::value;
                    const double c00 = pa[d] - su_ * pq[d];
                    const double c0p = qc[d] + st_ * pq[d];
                    g[d][0][0] = (d == 2) ? w[r] * ff : 1.0;
                    if constexpr (NMAX > 0) g[d][1][0] = c00 * g[d][0][0];
                    static_for<(NMAX > 1 ? NMAX - 1 : 0)>([&](auto NN) {
                        constexpr int n = decltype(NN)::value + 1;
                        g[d][n + 1][0] = fma(c00, g[d][n][0], (n * b10) * g[d][n - 1][0]);
                    });
                    static_for<MMAX>([&](auto MM) {
                        constexpr int m = decltype(MM)::value;
                        static_for<NMAX + 1>([&](auto NN) {
                            constexpr int n = decltype(NN)::value;
                            double v = c0p * g[d][n][m];
                            if constexpr (m > 0) v = fma(m * b01, g[d][n][m - 1], v);
                            if constexpr (n > 0) v = fma(n * b00, g[d][n - 1][m], v);
                            g[d][n][m + 1] = v;
                        });
                    });
                });
                static_for<NE>([&](auto EE) {
                    constexpr int e = decltype(EE)::value;
                    constexpr int le = cx_range_l(LA, e), ie = cx_range_i(LA, e);
                    constexpr int ex = cx_lx(le, ie), ey = cx_ly(le, ie), ez = cx_lz(le, ie);
                    static_for<NF>([&](auto FF) {
                        constexpr int f = decltype(FF)::value;
                        constexpr int lf = cx_range_l(LC, f), jf = cx_range_i(LC, f);
                        constexpr int fx = cx_lx(lf, jf), fy = cx_ly(lf, jf), fz = cx_lz(lf, jf);
                        const double xy = g[0][ex][fx] * g[1][ey][fy];
                        if constexpr (NCU == 1) {
                            if constexpr (NCT == 1) acc[e * NF + f] = fma(xy, g[2][ez][fz], acc[e * NF + f]);
                            else accu[e * NF + f] = fma(xy, g[2][ez][fz], accu[e * NF + f]);
                        } else {
                            val[e * NF + f] = fma(xy, g[2][ez][fz], val[e * NF + f]);
                        }
                    });
                });
            });
            };      // quad
            if constexpr (RS) {
                const double rr = fast_rsqrt(P.rs_w2 + a0);      // theta = omega^2 / (omega^2 + a0), src/g2e.c:4445
                const double th = P.rs_w2 * rr * rr;
                const double sq = P.rs_sign * rr;                // sign * |omega| * rr = sign * sqrt(theta)
#pragma unroll 1
                for (int pass = P.rs_pass0; pass < 2; pass++)
                    quad(pass ? x * th : x, pass ? fac * sq : fac, pass ? th : 1.0);
            } else {
                quad(x, fac, 1.0);
            }
            if constexpr (NCU > 1) {
#pragma unroll
                for (int c = 0; c < NCU; c++) {
                    const double cc = su[9 + c];
#pragma unroll
                    for (int i = 0; i < NEF; i++) {
                        if constexpr (NCT == 1) acc[c * NEF + i] = fma(cc, val[i], acc[c * NEF + i]);
                        else accu[c * NEF + i] = fma(cc, val[i], accu[c * NEF + i]);
                    }
                }
            }
        }
        if constexpr (NCT > 1) {
#pragma unroll
            for (int ct = 0; ct < NCT; ct++)
#pragma unroll
                for (int i = 0; i < NCU * NEF; i++) {
                    if constexpr (ACC_SMEM) s_acc[(ct * NCU * NEF + i) * REG_THREADS] = fma(ccT[ct], accu[i], s_acc[(ct * NCU * NEF + i) * REG_THREADS]);
                    else acc[ct * NCU * NEF + i] = fma(ccT[ct], accu[i], acc[ct * NCU * NEF + i]);
                }
        }
    }
    // --- epilogue: HRR, c2s, store ---
    // Stores go through a per-warp shared-memory transpose: a thread owns a whole (ab|cd) block, whose rows are
    // contiguous in the tile only inside one column, so direct stores would touch 32 different sectors per instruction
    // (measured: every single-primitive class saturated at ~900 GB/s of 8-byte sector writes).  Per column the 32 threads
    // stage their RB row values, then the warp writes the 32*RB values in tile order (runs of >= DA or DB doubles,
    // the whole RB when the pair is uncontracted; neighbouring quartets own neighbouring row blocks by construction).
    const int sa = P.tstride[to], sb = P.tstride[P.NTs + to];
    const long long sc = (long long)P.ustride[u] * P.ld, sd = (long long)P.ustride[P.NU_all + u] * P.ld;
    const int rowbase0 = (int)(P.trow[to] - P.row0);
    double *obase = P.out + P.ucol[u] * P.ld;
    const int nca_t = P.nca_t, nca_u = P.nca_u;
    // Fast flush (uncontracted T pairs whose 32 row blocks are adjacent -- the normal case thanks to the class-contiguous
    // row numbering): one column of the warp is ONE contiguous run of nact * RB doubles, so the values are staged in tile
    // order and written with plain lane + 32 k addressing; no gather table, no index arithmetic per element.
    bool fastflush = false;
    int nact = 0, rb0 = 0;
    if constexpr (NCT == 1) {
        const int rb_prev = __shfl_up_sync(0xffffffffu, rowbase0, 1);
        // my block itself must be RB contiguous rows (always true in the whole-job tiles; a dense shell-slice block
        // strides its second index by the slice's row count) and start where my left neighbour's ends
        const bool contiguous = (sa == 1 && (DB == 1 || sb == DA)) || (sb == 1 && (DA == 1 || sa == DB));
        fastflush = __all_sync(0xffffffffu, !active || (contiguous && (lane == 0 || rowbase0 == rb_prev + RB)));
        nact = __popc(__ballot_sync(0xffffffffu, active)) * RB;
        rb0 = __shfl_sync(0xffffffffu, rowbase0, 0);
    }
#pragma unroll 1
    for (int comb = 0; comb < NCT * NCU; comb++) {
        const int ct = comb / NCU, cu = comb - ct * NCU;
        if (cu == 0 && !fastflush) {
            // gather table of this T contraction block: element e = q * RB + r' of the warp's column slice, r' in tile order
            const int ca = ct % nca_t, cb = ct / nca_t;
            __syncwarp();
            // second field: > 0 -> a is the unit-stride index and the value is b's stride (single-shell pseudo pairs have
            // no b index and carry stride 0: any positive number will do); < 0 -> b is the unit-stride index, -value = a's stride
            s_meta[lane] = make_int2(active ? rowbase0 + ca * DA * sa + cb * DB * sb : -1, sa == 1 ? (sb > 0 ? sb : 1) : -sa);
            __syncwarp();
#pragma unroll
            for (int k = 0; k < RB; k++) {
                const int e = lane + 32 * k, q = e / RB, rp = e - q * RB;
                const int2 m = s_meta[q];
                int ma, mb, off;
                if (m.y > 0) { mb = rp / DA; ma = rp - mb * DA; off = ma + mb * m.y; }      // a is the first (unit-stride) index
                else { ma = rp / DB; mb = rp - ma * DB; off = mb - ma * m.y; }
                s_tab[e] = make_int2(m.x < 0 ? -1 : m.x + off, q * RB + ma * DB + mb);
            }
            __syncwarp();
        }
        // select the accumulator block of this combination (dynamic index -> predicated copies)
        double ef[NEF];
        if constexpr (NCT * NCU == 1) {
#pragma unroll
            for (int i = 0; i < NEF; i++) ef[i] = acc[i];
        } else if constexpr (ACC_SMEM) {
#pragma unroll
            for (int i = 0; i < NEF; i++) ef[i] = s_acc[(comb * NEF + i) * REG_THREADS];
        } else {
            static_for<NCT * NCU>([&](auto CI) {
                constexpr int ci = decltype(CI)::value;
                if (comb == ci) {
#pragma unroll
                    for (int i = 0; i < NEF; i++) ef[i] = acc[ci * NEF + i];
                }
            });
        }
        double abf[NFA * NFB * NF];
        hrr_pair_reg<LA, LB, 1, NF>(ef, abf, abT);                   // [ab][F]
        double abcd[NFA * NFB * NFC * NFD];
        hrr_pair_reg<LC, LD, NFA * NFB, 1>(abf, abcd, abU);          // [a][b][c][d]
        // c2s on each index with l >= 2 (s, p are identities)
        double s1[DA * NFB * NFC * NFD];
        if constexpr (LA >= 2 && !CART) c2s_reg<LA, 1, NFB * NFC * NFD>(abcd, s1);
        double *p1 = (LA >= 2 && !CART) ? s1 : abcd;
        double s2[DA * DB * NFC * NFD];
        if constexpr (LB >= 2 && !CART) c2s_reg<LB, DA, NFC * NFD>(p1, s2);
        double *p2 = (LB >= 2 && !CART) ? s2 : p1;
        double s3[DA * DB * DC * NFD];
        if constexpr (LC >= 2 && !CART) c2s_reg<LC, DA * DB, NFD>(p2, s3);
        double *p3 = (LC >= 2 && !CART) ? s3 : p2;
        double s4[DA * DB * DC * DD];
        if constexpr (LD >= 2 && !CART) c2s_reg<LD, DA * DB * DC, 1>(p3, s4);
        double *p4 = (LD >= 2 && !CART) ? s4 : p3;

        const int cc = cu % nca_u, cd = cu / nca_u;
        double *dst = obase + cc * DC * sc + cd * DD * sd;
        static_for<DD>([&](auto MD) {
            static_for<DC>([&](auto MC) {
                constexpr int mc = decltype(MC)::value, md = decltype(MD)::value;
                double *cp = dst + mc * sc + md * sd;
                if (fastflush) {
                    double *mine = s_st + lane * RB;
                    static_for<DA>([&](auto MA) {
                        static_for<DB>([&](auto MB) {
                            constexpr int ma = decltype(MA)::value, mb = decltype(MB)::value;
                            mine[ma * sa + mb * sb] = p4[((ma * DB + mb) * DC + mc) * DD + md];
                        });
                    });
                    __syncwarp();
                    double *run = cp + rb0 + lane;
#pragma unroll
                    for (int k = 0; k < RB; k++)
                        if (lane + 32 * k < nact) run[32 * k] = s_st[lane + 32 * k];
                    __syncwarp();
                } else {
                    static_for<DA>([&](auto MA) {
                        static_for<DB>([&](auto MB) {
                            constexpr int ma = decltype(MA)::value, mb = decltype(MB)::value;
                            s_st[lane * RB + ma * DB + mb] = p4[((ma * DB + mb) * DC + mc) * DD + md];
                        });
                    });
                    __syncwarp();
#pragma unroll
                    for (int k = 0; k < RB; k++) {
                        const int2 t = s_tab[lane + 32 * k];
                        if (t.x >= 0) cp[t.x] = s_st[t.y];
                    }
                    __syncwarp();
                }
            });
        });
    }
    }   // work items
    }   // batches
}
