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

// One HRR level on a [pre][part][post] array.  Input level holds, for every le in [l0, ltop],
// ncart(le) x ncart(jb-1) entries; output level holds le in [l0, ltop-1] with ncart(jb).
//   (a, b + 1_d | = (a + 1_d, b | + AB_d (a, b |
__device__ void hrr_level(const double *in, double *out, int pre, int post, int l0, int ltop, int jb,
                          const double *ab)
{
    const int nb_in = B200_NCART(jb - 1), nb_out = B200_NCART(jb);
    int in_part = 0, out_part = 0;
    for (int le = l0; le <= ltop; le++) in_part += B200_NCART(le) * nb_in;
    for (int le = l0; le < ltop; le++) out_part += B200_NCART(le) * nb_out;
    const int total = pre * out_part * post;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        int q = idx % post;
        int part = (idx / post) % out_part;
        int p = idx / (post * out_part);
        int le = l0, in_off = 0;
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
        const double *base = in + (size_t)p * in_part * post;
        double lo = base[(size_t)(in_off + ie * nb_in + ibp) * post + q];
        double hi = base[(size_t)(in_off + B200_NCART(le) * nb_in + iep * nb_in + ibp) * post + q];
        out[idx] = hi + ab[d] * lo;
    }
}

// out[p][m][q] = sum_c C[m][c] in[p][c][q]
__device__ void c2s_index(const double *in, double *out, int pre, int post, int l, const double *__restrict__ cmat)
{
    const int nin = B200_NCART(l), nout = 2 * l + 1;
    const int total = pre * nout * post;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        int q = idx % post;
        int m = (idx / post) % nout;
        int p = idx / (post * nout);
        const double *src = in + (size_t)p * nin * post + q;
        const double *cm = cmat + m * nin;
        double s = 0;
        for (int c = 0; c < nin; c++) s = fma(__ldg(cm + c), src[(size_t)c * post], s);
        out[idx] = s;
    }
}

__global__ void eri_generic_kernel(EngineParams P, GenericClass C, const Task *__restrict__ tasks, long long ntasks,
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

    double *s_rw = sm;                                   // [2*nroots]  t2,w interleaved as p
    double *s_g = s_rw + 2 * nreff;                      // [3][nreff][nmax+1][mmax+1]
    int *s_ecomp = (int *)(s_g + 3 * nreff * gstride_r);   // [nE], [nF] packed exponents
    int *s_fcomp = s_ecomp + nE;
    double *s_dyn = (double *)(s_fcomp + nF + ((nE + nF) & 1));
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

        for (int kq = 0; kq < (epi ? 0 : hk.npp); kq++) {
            const PrimPair pk = P.prims[hk.pp_off + kq];
            if (pk.cce > P.expcutoff) continue;
            for (int bq = 0; bq < hb.npp; bq++) {
                const PrimPair pb = P.prims[hb.pp_off + bq];
                if (pb.cce + pk.cce > P.expcutoff) continue;
                executed++;
                const double aij = pb.aij, akl = pk.aij;
                const double asum = aij + akl;
                const double a1 = aij * akl;
                const double a0 = a1 / asum;
                const double dx = pb.px - pk.px, dy = pb.py - pk.py, dz = pb.pz - pk.pz;
                double x = a0 * (dx * dx + dy * dy + dz * dz);
                double fac1 = common * pb.kij * pk.kij * sqrt(a0 / (a1 * a1 * a1));
                double theta = 1.0;
                const bool lr = P.omega > 0, sr = P.omega < 0;
                if (P.omega != 0) theta = P.omega * P.omega / (P.omega * P.omega + a0);
                if (lr) {                       // long-range attenuation, src/g2e.c:4477-4492
                    x *= theta;
                    fac1 *= sqrt(theta);
                }
                __syncthreads();                // previous primitive's G fully consumed
                if (tid < 2 * nroots) s_rw[tid] = rys_value(P.rys_coef, nroots, x, tid);
                else if (sr && tid < 4 * nroots) s_rw[tid] = rys_value(P.rys_coef, nroots, x * theta, tid - 2 * nroots);
                __syncthreads();
                if (tid < 3 * nreff) {
                    const int r = tid / 3, xyz = tid - 3 * r;
                    const bool second = r >= nroots;                       // long-range half of the SR rule
                    const double sc = (lr || second) ? theta : 1.0;
                    const double s = s_rw[2 * r] * sc;                      // t^2 (LR: theta t^2)
                    const double wgt = second ? -sqrt(theta) * s_rw[2 * r + 1] : s_rw[2 * r + 1];
                    const double sa = s * akl / asum, sk = s * aij / asum;
                    const double b00 = 0.5 * s / asum;
                    const double b10 = 0.5 * (1.0 - sa) / aij;
                    const double b01 = 0.5 * (1.0 - sk) / akl;
                    const double pq = xyz == 0 ? dx : (xyz == 1 ? dy : dz);
                    const double pa = (xyz == 0 ? pb.px : (xyz == 1 ? pb.py : pb.pz)) - hb.ra[xyz];
                    const double qc = (xyz == 0 ? pk.px : (xyz == 1 ? pk.py : pk.pz)) - hk.ra[xyz];
                    const double c00 = pa - sa * pq;
                    const double c0p = qc + sk * pq;
                    double *g = s_g + (size_t)(xyz * nreff + r) * gstride_r;
                    const int ms = mmax + 1;
                    g[0] = (xyz == 2) ? wgt * fac1 : 1.0;
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
                const double *ccb = P.pcoef + hb.cc_off + (size_t)bq * C.ncab;
                const double *cck = P.pcoef + hk.cc_off + (size_t)kq * C.nccd;
                for (int idx = tid; idx < nEF; idx += blockDim.x) {
                    const int e = idx / nF, f = idx - e * nF;
                    const int ec = s_ecomp[e], fc = s_fcomp[f];
                    const int ms = mmax + 1;
                    const int ox = (ec & 255) * ms + (fc & 255);
                    const int oy = ((ec >> 8) & 255) * ms + ((fc >> 8) & 255);
                    const int oz = ((ec >> 16) & 255) * ms + ((fc >> 16) & 255);
                    const double *gx = s_g + ox, *gy = s_g + (size_t)nreff * gstride_r + oy,
                                 *gz = s_g + (size_t)2 * nreff * gstride_r + oz;
                    double v = 0;
                    for (int r = 0; r < nreff; r++)
                        v = fma(gx[r * gstride_r] * gy[r * gstride_r], gz[r * gstride_r], v);
                    for (int ck = 0; ck < C.nccd; ck++) {
                        const double vk = v * __ldg(cck + ck);
                        for (int cb = 0; cb < C.ncab; cb++)
                            acc[(size_t)(ck * C.ncab + cb) * nEF + idx] += vk * __ldg(ccb + cb);
                    }
                }
            }
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
        const bool use_table = C.epi.rowptr != nullptr && cm == C.epi_cm;
        for (int comb = 0; comb < ncomb; comb++) {
            const int cab = comb % C.ncab, ccd = comb / C.ncab;
            const int ca = cab % hb.nca, cb = cab / hb.nca;
            const int cc = ccd % hk.nca, cd = ccd / hk.nca;
            const double *cur = acc + (size_t)comb * nEF;
            double *nxt = w0;
            if (use_table) {
                // table-driven epilogue: the stages are sparse maps precomputed per class (kernels.h:EpiTable)
                for (int sidx = 0; sidx < C.epi.nstages; sidx++) {
                    const EpiStage S = C.epi.st[sidx];
                    const double *ab = S.which == 1 ? hb.ab : hk.ab;
                    const int *rp = C.epi.rowptr + S.row0;
                    for (int idx = tid; idx < S.nout; idx += blockDim.x) {
                        double v = 0;
                        for (int e = rp[idx]; e < rp[idx + 1]; e++) {
                            const int2 en = C.epi.ent[e];
                            const double f = en.y ? C.epi.coef[e] * ab[en.y - 1] : C.epi.coef[e];
                            v = fma(f, cur[en.x], v);
                        }
                        nxt[idx] = v;
                    }
                    __syncthreads();
                    cur = nxt;
                    nxt = (nxt == w0) ? w1 : w0;
                }
                const int n_out = C.epi.n_out;
                double *dst = out + task.off + (long long)ca * da * task.sa + (long long)cb * db * task.sb
                            + (long long)cc * dc * task.sc + (long long)cd * dd * task.sd;
                const uchar4 *sidx4 = C.epi.store_idx + (task.sa <= task.sb ? 0 : n_out);
                for (int idx = tid; idx < n_out; idx += blockDim.x) {
                    const uchar4 m = sidx4[idx];
                    dst[(long long)m.x * task.sa + (long long)m.y * task.sb + m.z * task.sc + m.w * task.sd]
                        = cur[((m.x * db + m.y) * dc + m.z) * dd + m.w];
                }
                __syncthreads();
                continue;
            }
            for (int jb = 1; jb <= lb; jb++) {
                hrr_level(cur, nxt, 1, nF, la, la + lb - jb + 1, jb, hb.ab);
                __syncthreads();
                cur = nxt;
                nxt = (nxt == w0) ? w1 : w0;
            }
            for (int jd = 1; jd <= ld; jd++) {
                hrr_level(cur, nxt, nfa * nfb, 1, lc, lc + ld - jd + 1, jd, hk.ab);
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
            for (int idx = tid; idx < n_out; idx += blockDim.x) {
                int ma, mb, r;
                if (a_fast) { ma = idx % da; r = idx / da; mb = r % db; r /= db; }
                else        { mb = idx % db; r = idx / db; ma = r % da; r /= da; }
                const int mc = r % dc, md = r / dc;
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

// ---------------------------------------------------------------- table-driven epilogue (host side)
// Built once per (class, Cartesian mask 0 / 15) and device, cached for the life of the process: the exact index arithmetic of
// hrr_level / c2s_index / the store loop above, evaluated on the host.
#include <map>
#include <mutex>
#include <vector>
#include <array>
static void host_cart_xyz(int l, int idx, int *lx, int *ly, int *lz)
{
    int n = 0;
    for (int x = l; x >= 0; x--)
        for (int y = l - x; y >= 0; y--, n++)
            if (n == idx) { *lx = x; *ly = y; *lz = l - x - y; return; }
    *lx = *ly = *lz = 0;
}
static int host_cart_index(int lx, int lz, int l) { const int r = l - lx; return r * (r + 1) / 2 + lz; }

struct EpiHost { std::vector<int> rowptr; std::vector<int2> ent; std::vector<double> coef; std::vector<EpiStage> st; };

static void epi_add_hrr(EpiHost &H, int which, int pre, int post, int l0, int ltop, int jb)
{
    const int nb_in = B200_NCART(jb - 1), nb_out = B200_NCART(jb);
    int in_part = 0, out_part = 0;
    for (int le = l0; le <= ltop; le++) in_part += B200_NCART(le) * nb_in;
    for (int le = l0; le < ltop; le++) out_part += B200_NCART(le) * nb_out;
    const int total = pre * out_part * post;
    H.st.push_back(EpiStage{total, which, (int)H.rowptr.size()});
    for (int idx = 0; idx < total; idx++) {
        const int q = idx % post;
        int part = (idx / post) % out_part;
        const int p = idx / (post * out_part);
        int le = l0, in_off = 0;
        while (part >= B200_NCART(le) * nb_out) { part -= B200_NCART(le) * nb_out; in_off += B200_NCART(le) * nb_in; le++; }
        const int ie = part / nb_out, ib = part - ie * nb_out;
        int bx, by, bz, ax, ay, az;
        host_cart_xyz(jb, ib, &bx, &by, &bz);
        host_cart_xyz(le, ie, &ax, &ay, &az);
        const int d = bx ? 0 : (by ? 1 : 2);
        bx -= (d == 0); bz -= (d == 2);
        const int ibp = host_cart_index(bx, bz, jb - 1), iep = host_cart_index(ax + (d == 0), az + (d == 2), le + 1);
        const int base = p * in_part * post;
        H.rowptr.push_back((int)H.ent.size());
        H.ent.push_back(make_int2(base + (in_off + B200_NCART(le) * nb_in + iep * nb_in + ibp) * post + q, 0)); H.coef.push_back(1.0);
        H.ent.push_back(make_int2(base + (in_off + ie * nb_in + ibp) * post + q, d + 1)); H.coef.push_back(1.0);
    }
    H.rowptr.push_back((int)H.ent.size());
}

static void epi_add_c2s(EpiHost &H, int pre, int post, int l, const double *cmat)
{
    const int nin = B200_NCART(l), nout = 2 * l + 1, total = pre * nout * post;
    H.st.push_back(EpiStage{total, 1, (int)H.rowptr.size()});
    for (int idx = 0; idx < total; idx++) {
        const int q = idx % post, m = (idx / post) % nout, p = idx / (post * nout);
        H.rowptr.push_back((int)H.ent.size());
        for (int c = 0; c < nin; c++)
            if (cmat[m * nin + c] != 0.0) { H.ent.push_back(make_int2(p * nin * post + q + c * post, 0)); H.coef.push_back(cmat[m * nin + c]); }
    }
    H.rowptr.push_back((int)H.ent.size());
}

// c2s_host: the dense cart->sph matrices (host copy of what EngineParams::c2s holds), offsets c2s_off_table[l]
int epilogue_table(EpiTable *T, int la, int lb, int lc, int ld, int cart, const double *c2s_host, const int *c2s_off_table)
{
    struct Dev { EpiTable t; };
    static std::mutex mtx;
    static std::map<std::array<int, 6>, Dev> cache;
    std::lock_guard<std::mutex> lock(mtx);
    int dev = 0;
    cudaGetDevice(&dev);
    const std::array<int, 6> key = {dev, la, lb, lc, ld, cart};
    auto it = cache.find(key);
    if (it != cache.end()) { *T = it->second.t; return T->rowptr ? 0 : -1; }
    Dev D;
    memset(&D.t, 0, sizeof D.t);
    EpiHost H;
    const int nF = sum_ncart(lc, lc + ld);
    const int nfa = B200_NCART(la), nfb = B200_NCART(lb), nfc = B200_NCART(lc), nfd = B200_NCART(ld);
    const int da = (cart || la < 2) ? nfa : 2 * la + 1, db = (cart || lb < 2) ? nfb : 2 * lb + 1;
    const int dc = (cart || lc < 2) ? nfc : 2 * lc + 1, dd = (cart || ld < 2) ? nfd : 2 * ld + 1;
    for (int jb = 1; jb <= lb; jb++) epi_add_hrr(H, 1, 1, nF, la, la + lb - jb + 1, jb);
    for (int jd = 1; jd <= ld; jd++) epi_add_hrr(H, 2, nfa * nfb, 1, lc, lc + ld - jd + 1, jd);
    if (!cart) {
        if (la > 1) epi_add_c2s(H, 1, nfb * nfc * nfd, la, c2s_host + c2s_off_table[la]);
        if (lb > 1) epi_add_c2s(H, da, nfc * nfd, lb, c2s_host + c2s_off_table[lb]);
        if (lc > 1) epi_add_c2s(H, da * db, nfd, lc, c2s_host + c2s_off_table[lc]);
        if (ld > 1) epi_add_c2s(H, da * db * dc, 1, ld, c2s_host + c2s_off_table[ld]);
    }
    const int n_out = da * db * dc * dd;
    bool ok = H.st.size() <= 12 && H.ent.size() <= ((size_t)3 << 20) && da < 256 && db < 256 && dc < 256 && dd < 256;
    if (ok) {
        std::vector<uchar4> sidx(2 * (size_t)n_out);
        for (int fast = 0; fast < 2; fast++)
            for (int idx = 0; idx < n_out; idx++) {
                int ma, mb, r;
                if (fast == 0) { ma = idx % da; r = idx / da; mb = r % db; r /= db; }
                else           { mb = idx % db; r = idx / db; ma = r % da; r /= da; }
                sidx[(size_t)fast * n_out + idx] = make_uchar4((unsigned char)ma, (unsigned char)mb, (unsigned char)(r % dc), (unsigned char)(r / dc));
            }
        int *d_rp = nullptr; int2 *d_en = nullptr; double *d_cf = nullptr; uchar4 *d_si = nullptr;
        ok = cudaMalloc((void **)&d_rp, sizeof(int) * std::max<size_t>(1, H.rowptr.size())) == cudaSuccess &&
             cudaMalloc((void **)&d_en, sizeof(int2) * std::max<size_t>(1, H.ent.size())) == cudaSuccess &&
             cudaMalloc((void **)&d_cf, sizeof(double) * std::max<size_t>(1, H.coef.size())) == cudaSuccess &&
             cudaMalloc((void **)&d_si, sizeof(uchar4) * sidx.size()) == cudaSuccess &&
             cudaMemcpy(d_rp, H.rowptr.data(), sizeof(int) * H.rowptr.size(), cudaMemcpyHostToDevice) == cudaSuccess &&
             cudaMemcpy(d_en, H.ent.data(), sizeof(int2) * H.ent.size(), cudaMemcpyHostToDevice) == cudaSuccess &&
             cudaMemcpy(d_cf, H.coef.data(), sizeof(double) * H.coef.size(), cudaMemcpyHostToDevice) == cudaSuccess &&
             cudaMemcpy(d_si, sidx.data(), sizeof(uchar4) * sidx.size(), cudaMemcpyHostToDevice) == cudaSuccess;
        if (ok) {
            D.t.nstages = (int)H.st.size(); D.t.n_out = n_out;
            for (size_t k = 0; k < H.st.size(); k++) D.t.st[k] = H.st[k];
            D.t.rowptr = d_rp; D.t.ent = d_en; D.t.coef = d_cf; D.t.store_idx = d_si;
        } else { cudaFree(d_rp); cudaFree(d_en); cudaFree(d_cf); cudaFree(d_si); cudaGetLastError(); }
    }
    cache[key] = D;
    *T = D.t;
    return T->rowptr ? 0 : -1;
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

extern const double *engine_c2s_coef();
int epilogue_table(EpiTable *T, int la, int lb, int lc, int ld, int cart, const double *c2s_host, const int *c2s_off_table);

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
    size_t fixed = sizeof(double) * (2 * C->nreff + (size_t)3 * C->nreff * (nmax + 1) * (mmax + 1))
                 + sizeof(int) * (C->nE + C->nF + 2);
    size_t acc_b = sizeof(double) * nEF * ncab * nccd;
    size_t work_b = sizeof(double) * 2 * (size_t)C->work_size;
    const size_t budget = 96 * 1024;       // keeps >= 2 blocks per SM
    size_t smem = fixed;
    // table-driven epilogue (pure spherical / pure Cartesian output; needs a device: host-only planning skips it).  OFF by default:
    // measured slower than the run-time index decoding (C2H6 cc-pVQZ pass 143 -> 168 ms, (ff|ff) 2.2 -> 3.1 us per quartet) --
    // every block streams the class' whole table (~20 B per entry, up to 1 MB per quartet) from L2, which costs more than the
    // ~150 instructions per element it saves.  CINTB200_EPITAB=1 enables it (kept for the parity test of the maps).
    static const bool epi_on = getenv("CINTB200_EPITAB") && atoi(getenv("CINTB200_EPITAB"));
    int dev_probe = 0;
    C->epi_cm = cart ? 15 : 0;
    if (epi_on && (la + lb + lc + ld) > 0 && cudaGetDevice(&dev_probe) == cudaSuccess)
        epilogue_table(&C->epi, la, lb, lc, ld, cart, engine_c2s_coef(), c2s_off_table);
    else cudaGetLastError();
    static const bool wide_on = !(getenv("CINTB200_NO_WIDE") && atoi(getenv("CINTB200_NO_WIDE")));
    C->wide = wide_on && wide_eligible(la, lb, lc, ld, ncab, nccd, short_range);
    C->acc_in_smem = !C->wide && (smem + acc_b <= budget);      // wide classes: the accumulators arrive in global scratch
    if (C->acc_in_smem) smem += acc_b;
    C->work_in_smem = (smem + work_b <= budget);
    if (C->work_in_smem) smem += work_b;
    if (smem > 200 * 1024) return -1;
    C->scratch_per_block = (C->acc_in_smem ? 0 : nEF * ncab * nccd) + (C->work_in_smem ? 0 : 2 * (size_t)C->work_size);
    int threads = (int)((nEF + 31) / 32 * 32);
    if (threads < 96) threads = 96;        // >= 3 * nreff (<= 66 quadrature points x 3 axes handled by tid < 3*nreff)
    if (threads > 256) threads = 256;
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
    if (L.smem > 48 * 1024) {
        if (cudaFuncSetAttribute(eri_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem) != cudaSuccess)
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
            eri_generic_kernel<<<here, L.threads, L.smem, stream>>>(P, CE, tasks, ntasks, out, nonzero, counters, TP, uprefix);
            if (cudaGetLastError() != cudaSuccess) return -1;
        }
        return 0;
    }
    eri_generic_kernel<<<L.grid, L.threads, L.smem, stream>>>(P, C, tasks, ntasks, out, nonzero, counters, TP, uprefix);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
