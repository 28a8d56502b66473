// Cooperative ERI kernel: FS LANES PER SHELL QUARTET, for classes whose [e0|f0] block does not fit one
// thread's registers (d-rich classes of cc-pVDZ, anything with nEF > ~70).
//
// Work split inside a quartet (all compile-time shapes):
//   "register side" pair (LA LB|  : every lane keeps ALL NE components e of [e0| in registers;
//   "lane side"     pair |LC LD)  : lane f owns ONE component f of |f0]  (NF <= FS lanes busy).
// Per primitive quartet the FS lanes cooperate through shared memory:
//   lanes 0..2N-1   evaluate one root/weight polynomial each          (CINTrys_roots, src/rys_roots.c:57)
//   lanes 0..3N-1   run the 2-D VRR of one (root, axis) each          (CINTg0_2e_2d, src/g2e.c:272-421)
//   every lane      loads the three G columns of its f and adds       (CINTgout2e, src/cint2e.c:961)
//                   gx[ex] gy[ey] gz[ez] into its NE accumulators
// Epilogue: HRR + cart->sph of the register side per lane, one smem exchange, then HRR + cart->sph of the
// lane side with the lanes re-assigned to rows, and a store where consecutive lanes write consecutive rows
// (c2s_sph_2e1, src/cart2sph.c:5324).
// Either pair of the tile (T = per-quartet bra, U = block-uniform ket) can be the register side
// (REG_IS_T); the integral is symmetric under bra<->ket exchange, only the output strides swap.
#pragma once
#include "kern_reg.cuh"

// per-(root,axis) G block padded to an odd number of doubles so the 3N lanes that write them hit distinct banks
// resident blocks per SM the compiler must allow, by accumulator count per lane (measured per class on C60:
// small-accumulator classes gain 20-25% from 3-4 blocks, the (dd|x) classes with 62 accumulators lose from spills)
#ifndef COOP_THREADS
#define COOP_THREADS 128
#endif
#ifndef COOP_FUSED_ROOTS
#define COOP_FUSED_ROOTS 0      // 1: roots evaluated on the recurrence lanes; measured in round 2: C60 1204 vs 1156 ms, kept off
#endif
#ifndef COOP_MB16
#define COOP_MB16 4
#endif
#ifndef COOP_MB32
#define COOP_MB32 3
#endif
#ifndef COOP_MBX
#define COOP_MBX 2
#endif
__host__ __device__ constexpr int coop_min_blocks_default(int nacc)
{
    const int mb = (nacc <= 16 ? COOP_MB16 : nacc <= 32 ? COOP_MB32 : COOP_MBX) * 128 / COOP_THREADS;
    return mb > 0 ? mb : 1;
}
// per-class override (tune.inc), keyed by the kernel's own template arguments (register side first)
__host__ __device__ constexpr int coop_min_blocks(int la, int lb, int lc, int ld, int ncr, int ncl)
{
    int nr = 0;
    for (int l = la; l <= la + lb; l++) nr += (l + 1) * (l + 2) / 2;
    return tune_lookup(g_coop_tune, sizeof(g_coop_tune) / sizeof(ClassTune), tune_key(la, lb, lc, ld, ncr, ncl), false, coop_min_blocks_default(ncr * ncl * nr));
}

__host__ __device__ constexpr int coop_g_task(int nmax, int mmax) { return ((nmax + 1) * (mmax + 1)) | 1; }
__host__ __device__ constexpr int coop_g_size(int n, int nmax, int mmax) { return 3 * n * coop_g_task(nmax, mmax); }
__host__ __device__ constexpr int coop_xsz(int n, int nmax, int mmax, int nf, int nab)
{
    int a = (nf | 1) * nab, b = coop_g_size(n, nmax, mmax) + 2 * n;
    return (a > b ? a : b) | 1;
}

template <int LA, int LB, int LC, int LD, int NCR, int NCL, int FS, bool REG_IS_T, bool RS = false, bool CART = false>
__global__ void __launch_bounds__(COOP_THREADS, coop_min_blocks(LA, LB, LC, LD, NCR, NCL)) eri_coop_kernel(const TileParams P)
{
    constexpr int NMAX = LA + LB, MMAX = LC + LD;
    constexpr int N = (LA + LB + LC + LD) / 2 + 1;
    constexpr int NE = cx_nrange(LA, LA + LB), NF = cx_nrange(LC, LC + LD);
    constexpr int NFA = cx_ncart(LA), NFB = cx_ncart(LB), NFC = cx_ncart(LC), NFD = cx_ncart(LD);
    constexpr int DA = CART ? NFA : SphDim<LA>::value, DB = CART ? NFB : SphDim<LB>::value;          // CART: see kern_reg.cuh
    constexpr int DC = CART ? NFC : SphDim<LC>::value, DD = CART ? NFD : SphDim<LD>::value;
    constexpr int NAB = DA * DB;
    constexpr int QPB = COOP_THREADS / FS;                       // quartets per block
    constexpr int NCT = REG_IS_T ? NCR : NCL, NCU = REG_IS_T ? NCL : NCR;
    constexpr int USTR = 9 + NCU;
    constexpr int GSZ = coop_g_size(N, NMAX, MMAX);
    constexpr int GT = coop_g_task(NMAX, MMAX);
    constexpr int XSZ = coop_xsz(N, NMAX, MMAX, NF, NAB);       // per-quartet smem (G + roots, reused for the exchange)
    constexpr int MS = MMAX + 1;
    constexpr int NFP = NF | 1;                                 // odd row stride of the exchange buffer (bank conflicts)
    static_assert(NF <= FS, "lane side has more components than lanes");
    static_assert(FS <= 32 && (FS & (FS - 1)) == 0, "FS must be a power of two within a warp");

    extern __shared__ double smem[];
    const int tid = threadIdx.x;
    const int q = tid / FS, lane = tid % FS;

    // --- smem carve-up: Rys table | U primitives | per-quartet work areas; the table is staged once per block and the
    //     block walks a contiguous range of work items (one ket x QPB bras), see kern_reg.cuh ---
    double *s_rys = smem;
    const int nint = c_rys_meta.nint[N];
    {
        constexpr int ROW = (RYS_DEG + 1) * 2 * N;
        for (int i = tid; i < nint * ROW; i += COOP_THREADS) {
            int r = i / ROW, c = i - r * ROW;
            s_rys[r * rys_smem_stride(N) + c] = __ldg(P.rys + i);
        }
    }
    constexpr int NWARP = COOP_THREADS / 32, QPW = 32 / FS;     // quartets per warp
    const int warp = tid >> 5, wl = tid & 31;
    double *s_u = s_rys + nint * rys_smem_stride(N) + (size_t)warp * P.umax * USTR;        // this WARP's ket primitives [nppu <= umax]
    double *s_q = s_rys + nint * rys_smem_stride(N) + (size_t)NWARP * P.umax * USTR + (size_t)q * XSZ;      // this quartet's area
    double *s_rw = s_q + GSZ;                                   // [2N] t2/w of the current primitive
    const long long total = P.items ? P.nitems : (long long)P.gx * P.NU;
    int cur_by = -1, t_lo = P.t_begin, t_hi_k = P.t_end;
    PairHdr hu;
    __syncthreads();                    // table staged; warps are independent from here on (see kern_reg.cuh)
    for (;;) {
    long long item0 = 0;
    if (wl == 0) item0 = (long long)atomicAdd(P.counter, (unsigned int)P.batch);
    item0 = __shfl_sync(0xffffffffu, item0, 0);
    if (item0 >= total) break;
    const long long item1 = (item0 + P.batch < total) ? item0 + P.batch : total;
    for (long long item = item0; item < item1; item++) {
    int by, bx = 0, u, t_hi = P.t_end, t0l = 0;
    if (P.items) {                      // list mode, see kern_reg.cuh
        const int4 it = P.items[item];
        by = u = it.x; t0l = it.y; t_hi = it.y + it.z;
    } else {
        by = (int)(item / P.gx); bx = (int)(item - (long long)by * P.gx);
        u = P.u_first + P.u_step * by;
    }
    if (by != cur_by) {
        __syncwarp();
        hu = P.pairs[P.upair[u]];
        for (int i = wl; i < hu.npp; i += 32) {
            const PrimPair pp = P.prims[hu.pp_off + i];
            double *d = s_u + i * USTR;
            d[0] = pp.aij; d[1] = pp.inv_aij; d[2] = pp.px; d[3] = pp.py; d[4] = pp.pz;
            d[5] = pp.px - hu.ra[0]; d[6] = pp.py - hu.ra[1]; d[7] = pp.pz - hu.ra[2];
            d[8] = pp.kij;
            for (int c = 0; c < NCU; c++) d[9 + c] = P.pcoef[hu.cc_off + i * NCU + c];
        }
        __syncwarp();
        cur_by = by;
        t_lo = P.t_begin; t_hi_k = P.t_end;      // see kern_reg.cuh for the two tile orderings; once per ket, not per work item
        if (P.tri) {
            const int K = P.uK[u];
            if (P.tri == 1 || K >= P.tri_i0) {
                int lo = P.t_begin + P.tB, hi = P.t_end + P.tB;
                t_hi_k = hi;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (P.tI[mid] < K) lo = mid + 1; else hi = mid; }
                t_lo = lo;
            }
        }
    }
    if (!P.items) t_hi = t_hi_k;
    const int t0 = P.items ? t0l : t_lo + bx * QPW;
    if (t0 >= t_hi) continue;            // warp-uniform
    const int t = t0 + (q - warp * QPW);
    const bool active = t < t_hi;
    const int to = active ? t : t_hi - 1;                   // T occurrence / row of the pair table, see kern_reg.cuh
    const int tt = P.tsel ? P.tsel[to] : to;
    // Schwarz: if every quartet of this warp is bounded below the threshold, skip the primitive loops -- the
    // accumulators stay zero and the epilogue zero-fills the blocks (what the reference does for empty blocks)
    const bool negligible = P.schwarz_thr > 0 && P.tq[tt] * P.uq[u] < P.schwarz_thr;
    const int Qb = __all_sync(0xffffffffu, negligible) ? 0 : __reduce_max_sync(0xffffffffu, P.tnpp[tt]);

    const size_t NT = P.NT;
    double raT[3], abT[3];
#pragma unroll
    for (int d = 0; d < 3; d++) { raT[d] = P.tgeom[d * NT + tt]; abT[d] = P.tgeom[(3 + d) * NT + tt]; }
    const double abU[3] = {hu.ab[0], hu.ab[1], hu.ab[2]};
    constexpr double fsp0 = 0.282094791773878143, fsp1 = 0.488602511902919921;
    constexpr double common = 34.986836655249725693
        * (LA == 0 ? fsp0 : LA == 1 ? fsp1 : 1.0) * (LB == 0 ? fsp0 : LB == 1 ? fsp1 : 1.0)
        * (LC == 0 ? fsp0 : LC == 1 ? fsp1 : 1.0) * (LD == 0 ? fsp0 : LD == 1 ? fsp1 : 1.0);

    // my lane-side component
    const int fl = lane < NF ? lane : NF - 1;
    int fx, fy, fz;
    {
        int lf = LC, r = fl;
        while (r >= cx_ncart(lf)) { r -= cx_ncart(lf); lf++; }
        // decode (lx,ly,lz) of component r of degree lf
        int lx = lf, cnt = 0;
        while (r >= cnt + (lf - lx + 1)) { cnt += lf - lx + 1; lx--; }
        fx = lx; fy = (lf - lx) - (r - cnt); fz = r - cnt;
    }

    constexpr int NCOMB = NCR * NCL;
    double acc[NCOMB * NE];
#pragma unroll
    for (int i = 0; i < NCOMB * NE; i++) acc[i] = 0.0;

    const int nppu = hu.npp;
    for (int tq = 0; tq < Qb; tq++) {
        const size_t o = (size_t)tq * NT + tt;
        const size_t F = (size_t)P.Q * NT;
        const double aT = P.tprim[o], iaT = P.tprim[F + o];
        const double pT[3] = {P.tprim[2 * F + o], P.tprim[3 * F + o], P.tprim[4 * F + o]};
        const double kT = P.tprim[5 * F + o];
        double ccT[NCT];
#pragma unroll
        for (int c = 0; c < NCT; c++) ccT[c] = P.tprim[(6 + c) * F + o];
#pragma unroll 1
        for (int uq = 0; uq < nppu; uq++) {
            const double *su = s_u + uq * USTR;
            const double aU = su[0], iaU = su[1];
            const double asum = aT + aU;
            const double rs = fast_rsqrt(asum);
            const double inv = rs * rs;
            // register side = "bra" of the recurrences, lane side = "ket"
            const double aR = REG_IS_T ? aT : aU, aL = REG_IS_T ? aU : aT;
            const double iaR = REG_IS_T ? iaT : iaU, iaL = REG_IS_T ? iaU : iaT;
            double pq[3], pa[3], qc[3];
#pragma unroll
            for (int d = 0; d < 3; d++) {
                const double pu = su[2 + d], pt = pT[d];
                const double ta = pt - raT[d], uc = su[5 + d];
                pq[d] = REG_IS_T ? pt - pu : pu - pt;
                pa[d] = REG_IS_T ? ta : uc;
                qc[d] = REG_IS_T ? uc : ta;
            }
            const double a0 = aT * aU * inv;
            const double x = a0 * (pq[0] * pq[0] + pq[1] * pq[1] + pq[2] * pq[2]);
            const double fac0 = common * kT * su[8] * iaT * iaU * rs;
            // one contraction combination: the coefficient is folded into the z column and the products are added
            // straight into the accumulators (no per-primitive val[] block: NE registers and NE FMAs less)
            double val[NCOMB == 1 ? 1 : NE];
            if constexpr (NCOMB > 1) {
#pragma unroll
                for (int e = 0; e < NE; e++) val[e] = 0.0;
            }
            const double cc1 = (NCOMB == 1) ? ccT[0] * su[9] : 1.0;
            // range-separated Coulomb: pass 1 = attenuated rule (theta x, theta t^2, weight sign sqrt(theta)), pass 0 = full
            // Coulomb (skipped for the long-range operator); the plain operator runs the loop body once
            double th_ = 1.0, sq_ = 1.0;
            if constexpr (RS) { const double rr = fast_rsqrt(P.rs_w2 + a0); th_ = P.rs_w2 * rr * rr; sq_ = P.rs_sign * rr; }
#pragma unroll 1
            for (int pass = RS ? P.rs_pass0 : 1; pass < 2; pass++) {
            const bool att = RS && pass == 1;
            const double xq = att ? x * th_ : x, fac = att ? fac0 * sq_ : fac0, thp = att ? th_ : 1.0;
            // value of root/weight polynomial p (p = 2k: t_k^2, p = 2k + 1: w_k) at xq
            auto rys_eval = [&](const int p) -> double {
                if (xq >= 35.0 + 5.0 * N) {
                    const int k = N * (N - 1) / 2 + (p >> 1);
                    return (p & 1) ? c_rys_lx_v[k] * rsqrt(xq) : c_rys_lx_r[k] / xq;
                }
                int idx;
                double y;
                rys_locate(xq, idx, y);
                asm("" : "+r"(idx));        // opaque index: no 24-bit overflow of the folded grid offset in the LDS immediates
                const double *cf = s_rys + idx * rys_smem_stride(N) + p;
                // Estrin (depth 4) instead of Horner (depth 9): this phase is a dependent chain on a few busy lanes
                static_assert(RYS_DEG == 9, "Estrin scheme written for degree 9");
                const double y2 = y * y, y4 = y2 * y2, y8 = y4 * y4;
                const double p01 = fma(cf[1 * 2 * N], y, cf[0]), p23 = fma(cf[3 * 2 * N], y, cf[2 * 2 * N]);
                const double p45 = fma(cf[5 * 2 * N], y, cf[4 * 2 * N]), p67 = fma(cf[7 * 2 * N], y, cf[6 * 2 * N]);
                const double p89 = fma(cf[9 * 2 * N], y, cf[8 * 2 * N]);
                return fma(p89, y8, fma(fma(p67, y2, p45), y4, fma(p23, y2, p01)));
            };
#if !COOP_FUSED_ROOTS
            // --- roots: one polynomial per lane ---
            for (int p = lane; p < 2 * N; p += FS) s_rw[p] = rys_eval(p);
            __syncwarp();
#endif
            // --- VRR: one (root, axis) per lane ---
            const double rho_l = aL * inv, rho_r = aR * inv;
            for (int task = lane; task < 3 * N; task += FS) {
                const int r = task / 3, d = task - 3 * r;
#if COOP_FUSED_ROOTS
                // measured slower: every VRR lane evaluates the root it consumes, the z lane also the
                // weight -- no separate root phase, one smem round trip and one warp sync less per primitive
                const double t2r = rys_eval(2 * r);
                const double wr = (d == 2) ? rys_eval(2 * r + 1) : 0.0;
#else
                const double t2r = s_rw[2 * r];
#endif
                const double s = RS ? t2r * thp : t2r;
                const double sl = s * rho_l, sr = s * rho_r;
                const double b00 = 0.5 * s * inv;
                const double b10 = (0.5 - 0.5 * sl) * iaR;
                const double b01 = (0.5 - 0.5 * sr) * iaL;
                const double pqd = d == 0 ? pq[0] : d == 1 ? pq[1] : pq[2];
                const double pad = d == 0 ? pa[0] : d == 1 ? pa[1] : pa[2];
                const double qcd = d == 0 ? qc[0] : d == 1 ? qc[1] : qc[2];
                const double c00 = pad - sl * pqd;
                const double c0p = qcd + sr * pqd;
                double g[NMAX + 1][MMAX + 1];
#if COOP_FUSED_ROOTS
                g[0][0] = (d == 2) ? wr * fac : 1.0;
#else
                g[0][0] = (d == 2) ? s_rw[2 * r + 1] * fac : 1.0;
#endif
                if constexpr (NMAX > 0) g[1][0] = c00 * g[0][0];
                static_for<(NMAX > 1 ? NMAX - 1 : 0)>([&](auto NN) {
                    constexpr int n = decltype(NN)::value + 1;
                    g[n + 1][0] = fma(c00, g[n][0], (n * b10) * g[n - 1][0]);
                });
                static_for<MMAX>([&](auto MM) {
                    constexpr int m = decltype(MM)::value;
                    static_for<NMAX + 1>([&](auto NN) {
                        constexpr int n = decltype(NN)::value;
                        double v = c0p * g[n][m];
                        if constexpr (m > 0) v = fma(m * b01, g[n][m - 1], v);
                        if constexpr (n > 0) v = fma(n * b00, g[n - 1][m], v);
                        g[n][m + 1] = v;
                    });
                });
                double *dst = s_q + (size_t)task * GT;
                static_for<NMAX + 1>([&](auto NN) {
                    static_for<MMAX + 1>([&](auto MM) {
                        dst[decltype(NN)::value * MS + decltype(MM)::value] = g[decltype(NN)::value][decltype(MM)::value];
                    });
                });
            }
            __syncwarp();
            // --- quadrature sum for my f ---
            static_for<N>([&](auto RR) {
                constexpr int r = decltype(RR)::value;
                const double *gx = s_q + (size_t)(3 * r) * GT + fx;
                const double *gy = s_q + (size_t)(3 * r + 1) * GT + fy;
                const double *gz = s_q + (size_t)(3 * r + 2) * GT + fz;
                double cx[NMAX + 1], cy[NMAX + 1], cz[NMAX + 1];
#pragma unroll
                for (int n = 0; n <= NMAX; n++) { cx[n] = gx[n * MS]; cy[n] = gy[n * MS]; cz[n] = gz[n * MS]; }
                if constexpr (NCOMB == 1) {
#pragma unroll
                    for (int n = 0; n <= NMAX; n++) cz[n] *= cc1;
                }
                static_for<NE>([&](auto EE) {
                    constexpr int e = decltype(EE)::value;
                    constexpr int le = cx_range_l(LA, e), ie = cx_range_i(LA, e);
                    constexpr int ex = cx_lx(le, ie), ey = cx_ly(le, ie), ez = cx_lz(le, ie);
                    if constexpr (NCOMB == 1) acc[e] = fma(cx[ex] * cy[ey], cz[ez], acc[e]);
                    else val[e] = fma(cx[ex] * cy[ey], cz[ez], val[e]);
                });
            });
            }       // passes
            if constexpr (NCOMB == 1) {
            } else {
                static_for<NCL>([&](auto CL) {
                    static_for<NCR>([&](auto CR) {
                        constexpr int cl = decltype(CL)::value, cr = decltype(CR)::value;
                        const double cR = REG_IS_T ? ccT[cr] : su[9 + cr];
                        const double cL = REG_IS_T ? su[9 + cl] : ccT[cl];
                        const double cc = cR * cL;
#pragma unroll
                        for (int e = 0; e < NE; e++) acc[(cl * NCR + cr) * NE + e] = fma(cc, val[e], acc[(cl * NCR + cr) * NE + e]);
                    });
                });
            }
        }
    }
    __syncwarp();

    // --- epilogue ---
    // strides: (a,b) = register side, (c,d) = lane side, mapped onto rows (T) / columns (U) of the tile
    const int tsa = P.tstride[to], tsb = P.tstride[P.NTs + to];
    const long long usc = (long long)P.ustride[u] * P.ld, usd = (long long)P.ustride[P.NU_all + u] * P.ld;
    const long long s_a = REG_IS_T ? tsa : usc, s_b = REG_IS_T ? tsb : usd;
    const long long s_c = REG_IS_T ? usc : tsa, s_d = REG_IS_T ? usd : tsb;
    const int nca_r = REG_IS_T ? P.nca_t : P.nca_u, nca_l = REG_IS_T ? P.nca_u : P.nca_t;
    double abR[3], abL[3];
#pragma unroll
    for (int d = 0; d < 3; d++) { abR[d] = REG_IS_T ? abT[d] : abU[d]; abL[d] = REG_IS_T ? abU[d] : abT[d]; }
    double *obase = P.out + (P.trow[to] - P.row0) + P.ucol[u] * P.ld;
#pragma unroll 1
    for (int comb = 0; comb < NCOMB; comb++) {
        const int cl = comb / NCR, cr = comb - cl * NCR;
        double ef[NE];
        static_for<NCOMB>([&](auto CI) {
            constexpr int ci = decltype(CI)::value;
            if (comb == ci) {
#pragma unroll
                for (int i = 0; i < NE; i++) ef[i] = acc[ci * NE + i];
            }
        });
        // register side: HRR + c2s, all local
        double abc[NFA * NFB];
        hrr_pair_reg<LA, LB, 1, 1>(ef, abc, abR);
        double s1[DA * NFB];
        if constexpr (LA >= 2 && !CART) c2s_reg<LA, 1, NFB>(abc, s1);
        double *p1 = (LA >= 2 && !CART) ? s1 : abc;
        double s2[DA * DB];
        if constexpr (LB >= 2 && !CART) c2s_reg<LB, DA, 1>(p1, s2);
        double *p2 = (LB >= 2 && !CART) ? s2 : p1;
        // exchange: X[mab][f]
        if (lane < NF) {
#pragma unroll
            for (int i = 0; i < NAB; i++) s_q[i * NFP + lane] = p2[i];
        }
        __syncwarp();
        const int ca = cr % nca_r, cb = cr / nca_r, cc = cl % nca_l, cd = cl / nca_l;
        double *dst = obase + ca * DA * s_a + cb * DB * s_b + cc * DC * s_c + cd * DD * s_d;
        for (int mab = lane; mab < NAB; mab += FS) {
            double fr[NF];
#pragma unroll
            for (int f = 0; f < NF; f++) fr[f] = s_q[mab * NFP + f];
            double cdc[NFC * NFD];
            hrr_pair_reg<LC, LD, 1, 1>(fr, cdc, abL);
            double s3[DC * NFD];
            if constexpr (LC >= 2 && !CART) c2s_reg<LC, 1, NFD>(cdc, s3);
            double *p3 = (LC >= 2 && !CART) ? s3 : cdc;
            double s4[DC * DD];
            if constexpr (LD >= 2 && !CART) c2s_reg<LD, DC, 1>(p3, s4);
            double *p4 = (LD >= 2 && !CART) ? s4 : p3;
            const int ma = mab / DB, mb = mab - ma * DB;
            if (active) {
                double *d2 = dst + ma * s_a + mb * s_b;
                static_for<DC>([&](auto MC) {
                    static_for<DD>([&](auto MD) {
                        constexpr int mc = decltype(MC)::value, md = decltype(MD)::value;
                        d2[mc * s_c + md * s_d] = p4[mc * DD + md];
                    });
                });
            }
        }
        __syncwarp();
    }
    }   // work items
    }   // batches
}
