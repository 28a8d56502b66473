// Generic ERI kernel: one thread block per shell tuple, any angular momentum up to B200_LMAX, any
// contraction pattern, 4-centre (ij|kl) and 3-centre (ij|k) tuples, spherical or Cartesian output.
//
// It is the catch-all of the engine: classes without a specialised kernel (kern_spd.cu) run here.
// Stages per tuple, with the reference function each one replaces:
//   primitive loop + screening   CINT2e_loop            src/cint2e.c:660-758  (rule :694,:720)
//   roots / weights              CINTrys_roots          src/rys_roots.c:57     -> table polynomials (rys.cuh)
//   recurrence coefficients      CINTg0_2e              src/g2e.c:4518-4540    -> written in t^2
//   2-D VRR                      CINTg0_2e_2d           src/g2e.c:272-421      -> one thread per (root, axis), smem
//   [e0|f0] quadrature sum       CINTgout2e             src/cint2e.c:961       -> one thread per Cartesian element
//   contraction                  CINTprim_to_ctr_0/1    src/g1e.c:530-560      -> coefficient products from the pair table
//   HRR                          CINTg0_*2d_4d          src/g2e.c:428-693      -> applied ONCE per contracted tuple on
//                                                                                [e0|f0] (exact identity; the reference
//                                                                                applies it per primitive and root)
//   cart->sph, scatter           c2s_sph_2e1            src/cart2sph.c:5324    -> dense small matrices, strided store
#include <algorithm>
#include <cstdlib>
#include "types.h"
#include "rys.cuh"
#include "kernels.h"
#include "tile_task.cuh"

__constant__ int c_cart_off[2 * B200_LMAX + 2];                 // first component of degree l
__constant__ unsigned char c_cart_xyz[3 * 560];                 // (lx,ly,lz) of every component, l = 0..2*LMAX
// the same table in global memory for the lookups whose index differs from thread to thread (a divergent constant-memory
// read is serialised lane by lane; through L1 it is one transaction per distinct line)
__device__ unsigned char d_cart_xyz[3 * 560];

__device__ __forceinline__ int cart_index(int lx, int lz, int l)
{
    int r = l - lx;
    return r * (r + 1) / 2 + lz;
}

// Division of a small non-negative int by a block-uniform run-time divisor without the ~20-instruction software divide:
// q = (n * m) >> 32 with m = floor(2^32 / d) + 1, exact for n, d < 2^16 (every index of the epilogue is far below that...
// the products n * d stay < 2^32, which is the condition).
#define GEN_PRIM_DOUBLES 16             // per primitive quartet of a batch: valid, x, fac, theta, aij, akl, PQ[3], PA[3], QC[3]
struct FastDiv {
    unsigned long long m;
    int d;
    __device__ __forceinline__ explicit FastDiv(int d_) : m((0x100000000ull / (unsigned)d_) + 1), d(d_) {}
    __device__ __forceinline__ int div(int n) const { return (int)(((unsigned long long)(unsigned)n * m) >> 32); }
};

// One HRR level on a [pre][part][post] array.  Input level holds, for every le in [l0, ltop],
// ncart(le) x ncart(jb-1) entries; output level holds le in [l0, ltop-1] with ncart(jb).
//   (a, b + 1_d | = (a + 1_d, b | + AB_d (a, b |
// The map part -> (two source offsets inside the input slice, axis) is the same for every (pre, post) element: it is decoded
// once per level into shared memory (s_map, out_part ints), so an element costs two multiply-shift divisions, one map load,
// two loads, one FMA and one store instead of ~150 instructions of component decoding.
__device__ void hrr_level(const double *in, double *out, int pre, int post, int l0, int ltop, int jb,
                          const double *ab, int *s_map)
{
    const int nb_in = B200_NCART(jb - 1), nb_out = B200_NCART(jb);
    int in_part = 0, out_part = 0;
    for (int le = l0; le <= ltop; le++) in_part += B200_NCART(le) * nb_in;
    for (int le = l0; le < ltop; le++) out_part += B200_NCART(le) * nb_out;
    for (int part0 = threadIdx.x; part0 < out_part; part0 += blockDim.x) {
        int part = part0, le = l0, in_off = 0;
        while (part >= B200_NCART(le) * nb_out) {
            part -= B200_NCART(le) * nb_out;
            in_off += B200_NCART(le) * nb_in;
            le++;
        }
        int ie = part / nb_out, ib = part - ie * nb_out;
        const unsigned char *bc = d_cart_xyz + 3 * (jb * (jb + 1) * (jb + 2) / 6 + ib);        // first component of degree l: l(l+1)(l+2)/6
        const unsigned char *ac = d_cart_xyz + 3 * (le * (le + 1) * (le + 2) / 6 + ie);
        int bx = bc[0], by = bc[1], bz = bc[2];
        int ax = ac[0], az = ac[2];
        int d = bx ? 0 : (by ? 1 : 2);
        bx -= (d == 0); bz -= (d == 2);
        int ibp = cart_index(bx, bz, jb - 1);
        int iep = cart_index(ax + (d == 0), az + (d == 2), le + 1);
        const int lo = in_off + ie * nb_in + ibp, hi = in_off + B200_NCART(le) * nb_in + iep * nb_in + ibp;
        s_map[part0] = lo | (hi << 14) | (d << 28);           // offsets < 2^14: the largest level of l = 6 pairs has 1640 entries
    }
    __syncthreads();
    const double abx = ab[0], aby = ab[1], abz = ab[2];
    const int total = pre * out_part * post;
    const FastDiv dpost(post), dpart(out_part);
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int r = dpost.div(idx), q = idx - r * post;
        const int p = dpart.div(r), part = r - p * out_part;
        const int m = s_map[part];
        const int d = m >> 28;
        const double *base = in + (size_t)p * in_part * post + q;
        const double lo = base[(size_t)(m & 0x3fff) * post], hi = base[(size_t)((m >> 14) & 0x3fff) * post];
        out[idx] = fma(d == 0 ? abx : d == 1 ? aby : abz, lo, hi);
    }
}

// out[p][m][q] = sum_c C[m][c] in[p][c][q]; zero coefficients (more than half of every cart->sph matrix) are skipped
__device__ void c2s_index(const double *in, double *out, int pre, int post, int l, const double *__restrict__ cmat)
{
    const int nin = B200_NCART(l), nout = 2 * l + 1;
    const int total = pre * nout * post;
    const FastDiv dpost(post), dout(nout);
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int r = dpost.div(idx), q = idx - r * post;
        const int p = dout.div(r), m = r - p * nout;
        const double *src = in + (size_t)p * nin * post + q;
        const double *cm = cmat + m * nin;
        double s = 0;
        for (int c = 0; c < nin; c++) {
            const double cf = __ldg(cm + c);
            if (cf != 0.0) s = fma(cf, src[(size_t)c * post], s);
        }
        out[idx] = s;
    }
}

// Two instantiations: CAP128 = the compiler must fit two 256-thread blocks per SM (128 registers, 40 bytes of spills) -- the
// kernel is latency-bound at 3 resident 96-thread blocks per SM otherwise (174 registers).  Measured with the cap: C2H6 cc-pVQZ
// pass 98 -> 75 ms, (pp|ss) with 16 contraction combinations 0.36 -> 0.25 us per quartet, ERI share of the ip1 gradient loop
// 1.13 -> 1.00 s; the widest generally contracted classes ((ff|ff) x 16 combinations: +10 %) keep the uncapped build.
template <bool CAP128>
__global__ void __launch_bounds__(256, CAP128 ? 2 : 1) eri_generic_kernel(EngineParams P, GenericClass C, const Task *__restrict__ tasks, long long ntasks,
                                   double *__restrict__ out, int *__restrict__ nonzero, unsigned long long *counters,
                                   TileParams TP, const long long *__restrict__ uprefix)
{
    extern __shared__ double sm[];
    const int tid = threadIdx.x;
    const int la = C.la, lb = C.lb, lc = C.lc, ld = C.ld;
    const int nmax = la + lb, mmax = lc + ld, nroots = C.nroots;
    // short-range Coulomb (omega < 0): erfc = 1 - erf, evaluated as the full-Coulomb rule plus the long-range rule
    // with negated weights, i.e. 2*nroots quadrature points.  The reference does the same for order <= 3
    // (src/g2e.c:4455-4476) and switches to dedicated erfc roots above; the combined rule is exact for every order.
    const int nreff = C.nreff;
    const int nE = C.nE, nF = C.nF, nEF = nE * nF;
    const int ncomb = C.ncab * C.nccd;
    const int gstride_r = (nmax + 1) * (mmax + 1);

    double *s_prim = sm;                                 // [pbatch][GEN_PRIM_DOUBLES] scalars of the batch's primitive quartets
    double *s_rw = s_prim + C.pbatch * GEN_PRIM_DOUBLES; // [pbatch][2*nreff]  t2,w interleaved as p
    double *s_g = s_rw + C.pbatch * 2 * nreff;           // [pbatch][3][nreff][nmax+1][mmax+1]
    int *s_ecomp = (int *)(s_g + (size_t)C.pbatch * 3 * nreff * gstride_r);   // [nE], [nF] packed exponents
    int *s_fcomp = s_ecomp + nE;
    int *s_map = s_fcomp + nF;                            // [C.map_ints] HRR map of the current level
    double *s_dyn = (double *)(s_map + C.map_ints + ((nE + nF + C.map_ints) & 1));
    double *gscratch = C.scratch + (size_t)blockIdx.x * C.scratch_per_block;
    double *acc, *w0, *w1;
    if (C.acc_in_smem) { acc = s_dyn; s_dyn += (size_t)ncomb * nEF; }
    else { acc = gscratch; gscratch += (size_t)ncomb * nEF; }
    if (C.work_in_smem) { w0 = s_dyn; w1 = w0 + C.work_size; }
    else { w0 = gscratch; w1 = w0 + C.work_size; }

    // component tables of the [e0| and |f0] index ranges
    for (int e = tid; e < nE; e += blockDim.x) {
        int le = la, r = e;
        while (r >= B200_NCART(le)) { r -= B200_NCART(le); le++; }
        const unsigned char *c = c_cart_xyz + 3 * (c_cart_off[le] + r);
        s_ecomp[e] = c[0] | (c[1] << 8) | (c[2] << 16);
    }
    for (int f = tid; f < nF; f += blockDim.x) {
        int lf = lc, r = f;
        while (r >= B200_NCART(lf)) { r -= B200_NCART(lf); lf++; }
        const unsigned char *c = c_cart_xyz + 3 * (c_cart_off[lf] + r);
        s_fcomp[f] = c[0] | (c[1] << 8) | (c[2] << 16);
    }
    const double fsp[2] = {0.282094791773878143, 0.488602511902919921};
    const double common = 34.986836655249725693 /* 2 pi^3 / sqrt(pi) */
        * (la < 2 ? fsp[la] : 1.0) * (lb < 2 ? fsp[lb] : 1.0) * (lc < 2 ? fsp[lc] : 1.0) * (ld < 2 ? fsp[ld] : 1.0);

    // epilogue-only mode (C.epilogue_only): the [e0|f0] accumulators of task C.task_base + blockIdx.x were produced by the wide
    // kernel (kern_wide.cu) in this block's scratch; only HRR, cart->sph and the store remain
    const bool epi = C.epilogue_only != 0;
    for (long long t = epi ? C.task_base + blockIdx.x : blockIdx.x; t < ntasks; t += epi ? ntasks : gridDim.x) {
        const Task task = tasks ? tasks[t] : tile_task(TP, t);
        if (task.bra < 0) continue;          // block-uniform
        const PairHdr hb = P.pairs[task.bra];
        const PairHdr hk = P.pairs[task.ket];
        __syncthreads();
        if (!epi) for (int i = tid; i < ncomb * nEF; i += blockDim.x) acc[i] = 0.0;
        int executed = 0;

        // Primitive quartets in BATCHES of C.pbatch: the per-primitive phases (2N root polynomials, 3N recurrence columns) occupy a
        // handful of threads each, so a batch runs them for all its primitives at once and the block synchronises four times per
        // batch instead of three times per primitive quartet.  Order of summation is unchanged (flat index kq * npp_b + bq ascending).
        const int PB = C.pbatch;
        const int npq = epi ? 0 : hk.npp * hb.npp;
        const bool lr = P.omega > 0, sr = P.omega < 0;
        const FastDiv dF(nF), dnpb(hb.npp > 0 ? hb.npp : 1);
        for (int base = 0; base < npq; base += PB) {
            const int nb = min(PB, npq - base);
            // phase 0: scalars of every primitive quartet of the batch (one thread each)
            if (tid < nb) {
                const int pq = base + tid, kq = dnpb.div(pq), bq = pq - kq * hb.npp;
                const PrimPair pk = P.prims[hk.pp_off + kq];
                const PrimPair pb = P.prims[hb.pp_off + bq];
                double *sp = s_prim + tid * GEN_PRIM_DOUBLES;
                const bool ok = !(pk.cce > P.expcutoff) && !(pb.cce + pk.cce > P.expcutoff);
                const double aij = pb.aij, akl = pk.aij, asum = aij + akl, a1 = aij * akl, a0 = a1 / asum;
                const double dx = pb.px - pk.px, dy = pb.py - pk.py, dz = pb.pz - pk.pz;
                double x = a0 * (dx * dx + dy * dy + dz * dz);
                double fac1 = common * pb.kij * pk.kij * sqrt(a0 / (a1 * a1 * a1));
                double theta = 1.0;
                if (P.omega != 0) theta = P.omega * P.omega / (P.omega * P.omega + a0);
                if (lr) { x *= theta; fac1 *= sqrt(theta); }            // long-range attenuation, src/g2e.c:4477-4492
                sp[0] = ok ? 1.0 : 0.0; sp[1] = x; sp[2] = fac1; sp[3] = theta; sp[4] = aij; sp[5] = akl;
                sp[6] = dx; sp[7] = dy; sp[8] = dz;
                sp[9] = pb.px - hb.ra[0]; sp[10] = pb.py - hb.ra[1]; sp[11] = pb.pz - hb.ra[2];
                sp[12] = pk.px - hk.ra[0]; sp[13] = pk.py - hk.ra[1]; sp[14] = pk.pz - hk.ra[2];
            }
            __syncthreads();
            // phase 1: roots and weights, one (primitive, polynomial) per thread
            const int nrw = sr ? 4 * nroots : 2 * nroots;
            const FastDiv dnrw(nrw);
            for (int i = tid; i < nb * nrw; i += blockDim.x) {
                const int bi = dnrw.div(i), pidx = i - bi * nrw;
                const double *sp = s_prim + bi * GEN_PRIM_DOUBLES;
                if (sp[0] == 0.0) continue;
                s_rw[bi * 2 * nreff + pidx] = pidx < 2 * nroots ? rys_value(P.rys_coef, nroots, sp[1], pidx)
                                                                : rys_value(P.rys_coef, nroots, sp[1] * sp[3], pidx - 2 * nroots);
            }
            __syncthreads();
            // phase 2: recurrences, one (primitive, root, axis) per thread
            const int nvr = 3 * nreff;
            const FastDiv dnvr(nvr);
            for (int i = tid; i < nb * nvr; i += blockDim.x) {
                const int bi = dnvr.div(i), rx = i - bi * nvr;
                const double *sp = s_prim + bi * GEN_PRIM_DOUBLES;
                if (sp[0] == 0.0) continue;
                const int r = rx / 3, xyz = rx - 3 * r;
                const double theta = sp[3], aij = sp[4], akl = sp[5], asum = aij + akl;
                const bool second = r >= nroots;                       // long-range half of the SR rule
                const double sc = (lr || second) ? theta : 1.0;
                const double s = s_rw[bi * 2 * nreff + 2 * r] * sc;     // t^2 (LR: theta t^2)
                const double wgt = second ? -sqrt(theta) * s_rw[bi * 2 * nreff + 2 * r + 1] : s_rw[bi * 2 * nreff + 2 * r + 1];
                const double sa = s * akl / asum, sk = s * aij / asum;
                const double b00 = 0.5 * s / asum;
                const double b10 = 0.5 * (1.0 - sa) / aij;
                const double b01 = 0.5 * (1.0 - sk) / akl;
                const double pq = sp[6 + xyz], pa = sp[9 + xyz], qc = sp[12 + xyz];
                const double c00 = pa - sa * pq;
                const double c0p = qc + sk * pq;
                double *g = s_g + (size_t)bi * 3 * nreff * gstride_r + (size_t)(xyz * nreff + r) * gstride_r;
                const int ms = mmax + 1;
                g[0] = (xyz == 2) ? wgt * sp[2] : 1.0;
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
            // phase 3: quadrature sums and contraction, primitives of the batch in order.  Work item = ([e0|f0] component, group of
            // contraction combinations): blocks with fewer components than threads (s / p classes with general contractions)
            // split the ncomb accumulator updates of a component over CSPLIT threads instead of leaving most of the block idle
            const int csplit = C.csplit;
            const FastDiv dsplit(csplit);
            for (int item = tid; item < nEF * csplit; item += blockDim.x) {
                const int idx = dsplit.div(item), cg = item - idx * csplit;
                const int e = dF.div(idx), f = idx - e * nF;
                const int ec = s_ecomp[e], fc = s_fcomp[f];
                const int ms = mmax + 1;
                const int ox = (ec & 255) * ms + (fc & 255);
                const int oy = ((ec >> 8) & 255) * ms + ((fc >> 8) & 255);
                const int oz = ((ec >> 16) & 255) * ms + ((fc >> 16) & 255);
                for (int bi = 0; bi < nb; bi++) {
                    if (s_prim[bi * GEN_PRIM_DOUBLES] == 0.0) continue;
                    const double *gb = s_g + (size_t)bi * 3 * nreff * gstride_r;
                    const double *gx = gb + ox, *gy = gb + (size_t)nreff * gstride_r + oy, *gz = gb + (size_t)2 * nreff * gstride_r + oz;
                    double v = 0;
                    for (int r = 0; r < nreff; r++)
                        v = fma(gx[r * gstride_r] * gy[r * gstride_r], gz[r * gstride_r], v);
                    const int pq = base + bi, kq = dnpb.div(pq), bq = pq - kq * hb.npp;
                    const double *ccb = P.pcoef + hb.cc_off + (size_t)bq * C.ncab;
                    const double *cck = P.pcoef + hk.cc_off + (size_t)kq * C.nccd;
                    if (csplit == 1) {
                        for (int ck = 0; ck < C.nccd; ck++) {
                            const double vk = v * __ldg(cck + ck);
                            for (int cb = 0; cb < C.ncab; cb++)
                                acc[(size_t)(ck * C.ncab + cb) * nEF + idx] += vk * __ldg(ccb + cb);
                        }
                    } else {
                        for (int comb = cg; comb < ncomb; comb += csplit) {
                            const int ck = comb / C.ncab, cb = comb - ck * C.ncab;
                            acc[(size_t)comb * nEF + idx] += (v * __ldg(cck + ck)) * __ldg(ccb + cb);
                        }
                    }
                }
            }
            if (tid == 0)
                for (int bi = 0; bi < nb; bi++) executed += s_prim[bi * GEN_PRIM_DOUBLES] != 0.0;
            __syncthreads();                // the batch's scalars and G are consumed before the next batch overwrites them
        }
        __syncthreads();
        if (tid == 0 && !epi) {
            if (nonzero) nonzero[t] = executed > 0;
            if (counters) atomicAdd(counters, (unsigned long long)executed);
        }

        // ---- epilogue per contraction combination: HRR (bra, ket), c2s, strided store ----
        const int nfa = B200_NCART(la), nfb = B200_NCART(lb), nfc = B200_NCART(lc), nfd = B200_NCART(ld);
        const int cm = P.cart ? 15 : (task.flags & 15);          // indices that stay Cartesian (block-uniform)
        const int da = (cm & 1) ? nfa : 2 * la + 1, db = (cm & 2) ? nfb : 2 * lb + 1;
        const int dc = (cm & 4) ? nfc : 2 * lc + 1, dd = (cm & 8) ? nfd : 2 * ld + 1;
        for (int comb = 0; comb < ncomb; comb++) {
            const int cab = comb % C.ncab, ccd = comb / C.ncab;
            const int ca = cab % hb.nca, cb = cab / hb.nca;
            const int cc = ccd % hk.nca, cd = ccd / hk.nca;
            const double *cur = acc + (size_t)comb * nEF;
            double *nxt = w0;
            for (int jb = 1; jb <= lb; jb++) {
                hrr_level(cur, nxt, 1, nF, la, la + lb - jb + 1, jb, hb.ab, s_map);
                __syncthreads();
                cur = nxt;
                nxt = (nxt == w0) ? w1 : w0;
            }
            for (int jd = 1; jd <= ld; jd++) {
                hrr_level(cur, nxt, nfa * nfb, 1, lc, lc + ld - jd + 1, jd, hk.ab, s_map);
                __syncthreads();
                cur = nxt;
                nxt = (nxt == w0) ? w1 : w0;
            }
            if (cm != 15) {
                if (la > 1 && !(cm & 1)) { c2s_index(cur, nxt, 1, nfb * nfc * nfd, la, P.c2s + C.c2s_off[0]); __syncthreads(); cur = nxt; nxt = (nxt == w0) ? w1 : w0; }
                if (lb > 1 && !(cm & 2)) { c2s_index(cur, nxt, da, nfc * nfd, lb, P.c2s + C.c2s_off[1]); __syncthreads(); cur = nxt; nxt = (nxt == w0) ? w1 : w0; }
                if (lc > 1 && !(cm & 4)) { c2s_index(cur, nxt, da * db, nfd, lc, P.c2s + C.c2s_off[2]); __syncthreads(); cur = nxt; nxt = (nxt == w0) ? w1 : w0; }
                if (ld > 1 && !(cm & 8)) { c2s_index(cur, nxt, da * db * dc, 1, ld, P.c2s + C.c2s_off[3]); __syncthreads(); cur = nxt; nxt = (nxt == w0) ? w1 : w0; }
            }
            // store: thread index runs over the index with the smallest stride first
            const int n_out = da * db * dc * dd;
            double *dst = out + task.off + (long long)ca * da * task.sa + (long long)cb * db * task.sb
                        + (long long)cc * dc * task.sc + (long long)cd * dd * task.sd;
            const bool a_fast = task.sa <= task.sb;
            const FastDiv d1(a_fast ? da : db), d2(a_fast ? db : da), d3(dc);
            for (int idx = tid; idx < n_out; idx += blockDim.x) {
                int r = d1.div(idx);
                const int m1 = idx - r * d1.d;
                const int r2 = d2.div(r), m2 = r - r2 * d2.d;
                const int ma = a_fast ? m1 : m2, mb = a_fast ? m2 : m1;
                const int md = d3.div(r2), mc = r2 - md * dc;
                dst[(long long)ma * task.sa + (long long)mb * task.sb + mc * task.sc + md * task.sd]
                    = cur[((ma * db + mb) * dc + mc) * dd + md];
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------- host side
static int sum_ncart(int l0, int l1) { int s = 0; for (int l = l0; l <= l1; l++) s += B200_NCART(l); return s; }

// sizes of every intermediate of the epilogue, to dimension the ping-pong buffers
static size_t epilogue_work_size(int la, int lb, int lc, int ld, int cart)
{
    size_t mx = 1;
    const int nF = sum_ncart(lc, lc + ld);
    const int nfa = B200_NCART(la), nfb = B200_NCART(lb), nfc = B200_NCART(lc), nfd = B200_NCART(ld);
    for (int jb = 1; jb <= lb; jb++) {
        size_t part = 0;
        for (int le = la; le <= la + lb - jb; le++) part += (size_t)B200_NCART(le) * B200_NCART(jb);
        if (part * nF > mx) mx = part * nF;
    }
    for (int jd = 1; jd <= ld; jd++) {
        size_t part = 0;
        for (int lf = lc; lf <= lc + ld - jd; lf++) part += (size_t)B200_NCART(lf) * B200_NCART(jd);
        if (part * nfa * nfb > mx) mx = part * nfa * nfb;
    }
    size_t full = (size_t)nfa * nfb * nfc * nfd;
    if (full > mx) mx = full;
    (void)cart;
    return mx;
}

int generic_setup_constants()
{
    int off[2 * B200_LMAX + 2];
    static unsigned char xyz[3 * 560];
    int n = 0;
    for (int l = 0; l <= 2 * B200_LMAX; l++) {
        off[l] = n;
        for (int lx = l; lx >= 0; lx--)
            for (int ly = l - lx; ly >= 0; ly--, n++) {
                xyz[3 * n] = (unsigned char)lx;
                xyz[3 * n + 1] = (unsigned char)ly;
                xyz[3 * n + 2] = (unsigned char)(l - lx - ly);
            }
    }
    off[2 * B200_LMAX + 1] = n;
    if (n > 560) return -1;
    if (cudaMemcpyToSymbol(c_cart_off, off, sizeof off) != cudaSuccess) return -1;
    if (cudaMemcpyToSymbol(c_cart_xyz, xyz, 3 * n) != cudaSuccess) return -1;
    if (cudaMemcpyToSymbol(d_cart_xyz, xyz, 3 * n) != cudaSuccess) return -1;
    return 0;
}

// Plan a launch for one class; returns 0 on success.  scratch is (re)allocated by the caller.
int generic_plan(GenericClass *C, GenericLaunch *L, int la, int lb, int lc, int ld, int ncab, int nccd,
                 int cart, long long ntasks, const int *c2s_off_table, int short_range)
{
    memset(C, 0, sizeof *C);
    C->la = la; C->lb = lb; C->lc = lc; C->ld = ld;
    C->nroots = (la + lb + lc + ld) / 2 + 1;
    if (C->nroots > RYS_NMAX) return -1;
    C->nreff = short_range ? 2 * C->nroots : C->nroots;
    C->ncab = ncab; C->nccd = nccd;
    C->nE = sum_ncart(la, la + lb);
    C->nF = sum_ncart(lc, lc + ld);
    C->work_size = (int)epilogue_work_size(la, lb, lc, ld, cart);
    C->c2s_off[0] = c2s_off_table[la]; C->c2s_off[1] = c2s_off_table[lb];
    C->c2s_off[2] = c2s_off_table[lc]; C->c2s_off[3] = c2s_off_table[ld];
    const size_t nEF = (size_t)C->nE * C->nF;
    const int nmax = la + lb, mmax = lc + ld;
    // largest HRR level (entries of one [part] slice) of either pair: the shared-memory map of hrr_level
    C->map_ints = 0;
    for (int side = 0; side < 2; side++) {
        const int l0 = side ? lc : la, l1 = side ? ld : lb;
        for (int j = 1; j <= l1; j++) {
            int part = 0;
            for (int le = l0; le <= l0 + l1 - j; le++) part += B200_NCART(le) * B200_NCART(j);
            if (part > C->map_ints) C->map_ints = part;
        }
    }
    // FastDiv (kernel) is exact while index x divisor < 2^32: the largest stage times the largest pass-through extent
    {
        const double post_max = std::max<double>(C->nF, (double)B200_NCART(lb) * B200_NCART(lc) * B200_NCART(ld));
        if ((double)C->work_size * post_max >= 4294967296.0) return -1;
    }
    // primitive quartets per batch: as many as ~24 kB of per-primitive state (scalars, roots, G) allow, at most 16
    const size_t per_prim = sizeof(double) * (GEN_PRIM_DOUBLES + 2 * C->nreff + (size_t)3 * C->nreff * (nmax + 1) * (mmax + 1));
    C->pbatch = (int)std::min<size_t>(16, std::max<size_t>(1, (24 * 1024) / per_prim));
    size_t fixed = per_prim * C->pbatch + sizeof(int) * (C->nE + C->nF + C->map_ints + 2);
    size_t acc_b = sizeof(double) * nEF * ncab * nccd;
    size_t work_b = sizeof(double) * 2 * (size_t)C->work_size;
    const size_t budget = 96 * 1024;       // keeps >= 2 blocks per SM
    size_t smem = fixed;
    static const bool wide_on = !(getenv("CINTB200_NO_WIDE") && atoi(getenv("CINTB200_NO_WIDE")));
    C->wide = wide_on && wide_eligible(la, lb, lc, ld, ncab, nccd, short_range);
    C->acc_in_smem = !C->wide && (smem + acc_b <= budget);      // wide classes: the accumulators arrive in global scratch
    if (C->acc_in_smem) smem += acc_b;
    C->work_in_smem = (smem + work_b <= budget);
    if (C->work_in_smem) smem += work_b;
    if (smem > 200 * 1024) return -1;
    C->scratch_per_block = (C->acc_in_smem ? 0 : nEF * ncab * nccd) + (C->work_in_smem ? 0 : 2 * (size_t)C->work_size);
    int threads = (int)((nEF + 31) / 32 * 32);          // measured: sizing the block for nEF x ncomb items costs more in resident blocks than it gains
    if (threads < 96) threads = 96;
    if (threads > 256) threads = 256;
    // contraction combinations of one component spread over csplit threads when the block has threads to spare
    C->csplit = 1;
    while (C->csplit * 2 <= ncab * nccd && nEF * (size_t)(C->csplit * 2) <= (size_t)threads) C->csplit *= 2;
    L->threads = threads;
    L->smem = smem;
    int per_sm = (int)(budget * 2 / (smem > 4096 ? smem : 4096));
    if (per_sm > 2048 / threads) per_sm = 2048 / threads;
    if (per_sm > 16) per_sm = 16;
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)148 * per_sm;
    if (C->wide) {
        // wide classes: one block per task and launch pair (quadrature, epilogue); as many tasks per pair as 2 GB of scratch
        // hold, so that a class is a handful of full-machine launches instead of a chain of single-wave ones
        const long long cap = (long long)(((size_t)2 << 30) / (sizeof(double) * std::max<size_t>(1, C->scratch_per_block)));
        grid = std::max<long long>(grid, std::min<long long>(cap, 1 << 20));
    }
    if (grid > ntasks) grid = ntasks;
    L->grid = (int)grid;
    return 0;
}

int generic_launch(const EngineParams &P, const GenericClass &C, const GenericLaunch &L, const Task *tasks,
                   long long ntasks, double *out, int *nonzero, unsigned long long *counters, cudaStream_t stream,
                   const TileParams *tile, const long long *uprefix)
{
    TileParams TP;
    memset(&TP, 0, sizeof TP);
    if (tile) TP = *tile;
    if (ntasks <= 0) return 0;
    // register-capped build unless the class is one of the widest contracted ones (threshold between (dd|dd) and (ff|ff))
    const bool cap = (size_t)C.nE * C.nF < 3000 || C.ncab * C.nccd == 1;
    auto kernel = cap ? eri_generic_kernel<true> : eri_generic_kernel<false>;
    if (L.smem > 48 * 1024) {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem) != cudaSuccess)
            return -1;
    }
    if (C.wide) {
        // high-l classes: quadrature by the wide kernel into this launch's scratch blocks (L.grid tasks at a time), then the
        // epilogue of this kernel on that scratch
        GenericClass CE = C;
        CE.epilogue_only = 1;
        for (long long base = 0; base < ntasks; base += L.grid) {
            const int here = (int)std::min<long long>(L.grid, ntasks - base);
            if (wide_launch(P, C, tasks, base, ntasks, here, nonzero, stream, tile)) return -1;
            CE.task_base = base;
            kernel<<<here, L.threads, L.smem, stream>>>(P, CE, tasks, ntasks, out, nonzero, counters, TP, uprefix);
            if (cudaGetLastError() != cudaSuccess) return -1;
        }
        return 0;
    }
    kernel<<<L.grid, L.threads, L.smem, stream>>>(P, C, tasks, ntasks, out, nonzero, counters, TP, uprefix);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
