/*
 * TEST INFRASTRUCTURE (CPU oracle) -- never linked into or called from the product path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this file.
 *
 * Plain-C restatement of the reference's Rys-quadrature ERI path, one shell tuple per call:
 *   int2e_sph / int2e_cart      src/cint2e.c:1186,1202 -> CINT2e_drv src/cint2e.c:794
 *   int3c2e_sph / int3c2e_cart  src/cint3c2e.c:693     -> CINT3c2e_drv src/cint3c2e.c:556
 * Stages, each citing what it follows:
 *   quartet constants      CINTinit_int2e_EnvVars  src/g2e.c:21-172, CINTinit_int3c2e_EnvVars src/g3c2e.c:21
 *   pair screening (cce)   CINTset_pairdata        src/optimizer.c:288-342
 *   primitive loop         CINT2e_loop_nopt        src/cint2e.c:57-262 (screening rules :204,:227)
 *   g0 / coefficients      CINTg0_2e               src/g2e.c:4425-4545 (incl. LR/SR switches :4443-4492)
 *   2-D VRR                CINTg0_2e_2d            src/g2e.c:272-421
 *   4-D HRR                CINTg0_{lj,kj,il,ik}2d_4d src/g2e.c:428-693
 *   gout                   CINTgout2e              src/cint2e.c:961
 *   contraction            CINTprim_to_ctr_0/1     src/g1e.c:530-560
 *   cart->sph + scatter    c2s_sph_2e1             src/cart2sph.c:5324, c2s_sph_3c2e1 :5884, c2s_cart_2e1 :5845
 * Deliberate differences (results agree to rounding, not bit-for-bit):
 *   - Rys roots come from oracle/rys_stieltjes.c (extended-precision Golub-Welsch), which is MORE
 *     accurate than the reference's fits for nroots >= 6 (see tests/test_oracle.py tolerances);
 *   - contraction is a direct 4-index accumulation instead of the staged gctri/j/k/l buffers;
 *   - c2s is a dense matrix product with coefficients computed here from the closed form of the
 *     real solid harmonics (the reference hard-codes sparse forms for d,f,g).
 * Parity status: PINNED (tests/test_oracle.py: reference known answers + element-wise vs oracle/_ref).
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "eri_oracle.h"

#define ATM_SLOTS 6
#define BAS_SLOTS 8
#define ATOM_OF 0
#define ANG_OF 1
#define NPRIM_OF 2
#define NCTR_OF 3
#define PTR_EXP 5
#define PTR_COEFF 6
#define PTR_COORD 1
#define PTR_EXPCUTOFF 0
#define PTR_RANGE_OMEGA 8
#define LMAX_ORACLE 10
#define MAXROOTS 32

static int ncart(int l) { return (l + 1) * (l + 2) / 2; }

/* src/g1e.c:565-572 */
static double fac_sp(int l)
{
        if (l == 0) return 0.282094791773878143;
        if (l == 1) return 0.488602511902919921;
        return 1;
}

/* src/misc.c:86 -- 1/sqrt(int r^(2l+2) exp(-2 a r^2) dr) */
double oracle_gto_norm(int l, double a)
{
        double p = l + 1.5;
        return 1.0 / sqrt(tgamma(p) / (2.0 * pow(2.0 * a, p)));
}

/* src/cint_bas.c:166-181: lx descending, then ly descending */
static void cart_comp(int l, int *nx, int *ny, int *nz)
{
        int lx, ly, n = 0;
        for (lx = l; lx >= 0; lx--)
                for (ly = l - lx; ly >= 0; ly--, n++) {
                        nx[n] = lx;
                        ny[n] = ly;
                        nz[n] = l - lx - ly;
                }
}

static long double lfact(int n) { long double r = 1; while (n > 1) r *= n--; return r; }
static long double lbinom(int n, int k) { return (k < 0 || k > n) ? 0 : lfact(n) / (lfact(k) * lfact(n - k)); }

/*
 * Real solid harmonics in Cartesian monomials (rows m=-l..l, libcint ordering; l<=1 identity because
 * the reference folds the s,p constants into common_factor: src/g2e.c:54-56).  Closed form:
 *   S_l^{|m|} ~ sum_{t,u,v} ... the usual expansion of r^l P_l^|m|(cos th) {cos,sin}(|m| phi).
 */
int oracle_c2s_matrix(int l, double *c2s)
{
        int nc = ncart(l), ns = 2 * l + 1;
        int nx[128], ny[128], nz[128];
        int m, k, p, a, b, n;
        if (l > LMAX_ORACLE) return 1;
        memset(c2s, 0, sizeof(double) * ns * nc);
        if (l <= 1) {
                for (n = 0; n < nc; n++) c2s[n * nc + n] = 1;
                return 0;
        }
        cart_comp(l, nx, ny, nz);
        for (m = -l; m <= l; m++) {
                int am = abs(m);
                long double norm = sqrtl((2 * l + 1) / (4 * 3.141592653589793238462643383279502884L));
                if (am) norm *= sqrtl(2 * lfact(l - am) / lfact(l + am));
                double *row = c2s + (m + l) * nc;
                for (k = 0; k <= (l - am) / 2; k++) {
                        /* Pi_l^m term: coef * r2^k * z^(l-2k-am) */
                        long double ck = ((k & 1) ? -1.0L : 1.0L) * lbinom(l, k) * lbinom(2 * l - 2 * k, l)
                                * lfact(l - 2 * k) / (powl(2, l) * lfact(l - 2 * k - am));
                        /* r2^k = sum_{a+b<=k} k!/(a! b! (k-a-b)!) x^2a y^2b z^2(k-a-b) */
                        for (a = 0; a <= k; a++)
                        for (b = 0; b <= k - a; b++) {
                                long double cm = lfact(k) / (lfact(a) * lfact(b) * lfact(k - a - b));
                                /* (x+iy)^am = sum_p C(am,p) x^p (iy)^(am-p) */
                                for (p = 0; p <= am; p++) {
                                        int q = am - p;
                                        int phase;
                                        if (m >= 0) phase = (q % 4 == 0) ? 1 : (q % 4 == 2) ? -1 : 0;
                                        else        phase = (q % 4 == 1) ? 1 : (q % 4 == 3) ? -1 : 0;
                                        if (!phase) continue;
                                        int ex = 2 * a + p, ey = 2 * b + q, ez = 2 * (k - a - b) + l - 2 * k - am;
                                        for (n = 0; n < nc; n++)
                                                if (nx[n] == ex && ny[n] == ey && nz[n] == ez)
                                                        row[n] += (double)(norm * ck * cm * lbinom(am, p) * phase);
                                }
                        }
                }
        }
        return 0;
}

typedef struct {
        int l, nprim, nctr;
        const double *r, *a, *c;        /* centre, exponents, coeff[nctr][nprim] */
        double logmaxc[64];             /* CINTOpt_log_max_pgto_coeff src/optimizer.c:259 */
} Shell;

static void load_shell(Shell *s, int ish, const int *atm, const int *bas, const double *env)
{
        int ip, ic;
        s->l = bas[ish * BAS_SLOTS + ANG_OF];
        s->nprim = bas[ish * BAS_SLOTS + NPRIM_OF];
        s->nctr = bas[ish * BAS_SLOTS + NCTR_OF];
        s->r = env + atm[bas[ish * BAS_SLOTS + ATOM_OF] * ATM_SLOTS + PTR_COORD];
        s->a = env + bas[ish * BAS_SLOTS + PTR_EXP];
        s->c = env + bas[ish * BAS_SLOTS + PTR_COEFF];
        for (ip = 0; ip < s->nprim && ip < 64; ip++) {
                double mx = 0;
                for (ic = 0; ic < s->nctr; ic++) {
                        double v = fabs(s->c[ic * s->nprim + ip]);
                        if (v > mx) mx = v;
                }
                s->logmaxc[ip] = log(mx);
        }
}

/* g index helper for the 4-D array g[xyz][root][i][j][k][l] */
typedef struct { int nr, di, dj, dk, dl; } GDim;
static size_t gidx(const GDim *d, int r, int i, int j, int k, int l)
{
        return ((((size_t)l * d->dk + k) * d->dj + j) * d->di + i) * d->nr + r;
}

/*
 * One shell tuple.  ncenter = 4: (ij|kl); ncenter = 3: (ij|k) with the auxiliary shell on the ket
 * and a fictitious l-centre with exponent 0 (src/g3c2e.c:66-105).
 */
static int eri_tuple(double *out, const int *dims, const int *shls, int ncenter, int sph,
                     const int *atm, int natm, const int *bas, int nbas, const double *env)
{
        Shell sh[4];
        static const double zero_exp[1] = {0.0}, one_coef[1] = {1.0};
        int n, ic;
        (void)natm; (void)nbas;
        if (ncenter == 2) {
                /* (i|k): aj = al = 0, rij = ri, rkl = rk (src/g2c2e.c:15-100); no primitive screening (src/cint2c2e.c:113-150) */
                load_shell(&sh[0], shls[0], atm, bas, env);
                load_shell(&sh[2], shls[1], atm, bas, env);
                sh[1].l = 0; sh[1].nprim = 1; sh[1].nctr = 1;
                sh[1].r = sh[0].r; sh[1].a = zero_exp; sh[1].c = one_coef; sh[1].logmaxc[0] = 0;
        } else {
                for (n = 0; n < ncenter; n++) load_shell(&sh[n], shls[n], atm, bas, env);
        }
        if (ncenter <= 3) {
                sh[3].l = 0; sh[3].nprim = 1; sh[3].nctr = 1;
                sh[3].r = sh[2].r; sh[3].a = zero_exp; sh[3].c = one_coef; sh[3].logmaxc[0] = 0;
        }
        const int li = sh[0].l, lj = sh[1].l, lk = sh[2].l, ll = sh[3].l;
        const int nfi = ncart(li), nfj = ncart(lj), nfk = ncart(lk), nfl = ncart(ll);
        const int nf = nfi * nfj * nfk * nfl;
        const int nci = sh[0].nctr, ncj = sh[1].nctr, nck = sh[2].nctr, ncl = sh[3].nctr;
        const int nc = nci * ncj * nck * ncl;
        const double omega = env[PTR_RANGE_OMEGA];

        double common = (M_PI * M_PI * M_PI) * 2 / 1.7724538509055160272981674833411451
                * fac_sp(li) * (ncenter == 2 ? 1.0 : fac_sp(lj)) * fac_sp(lk);
        double expcutoff;
        if (ncenter == 2) {
                expcutoff = 1e300;
        } else if (ncenter == 4) {
                common *= fac_sp(ll);
                expcutoff = (env[PTR_EXPCUTOFF] == 0) ? 60 : fmax(40, env[PTR_EXPCUTOFF]) + 1;
        } else {
                expcutoff = (env[PTR_EXPCUTOFF] == 0) ? 60 : fmax(40, env[PTR_EXPCUTOFF]);
        }
        const int order = (li + lj + lk + ll) / 2 + 1;
        const int nroots = (omega < 0 && order <= 3) ? 2 * order : order;

        /* output block geometry */
        const int di = (sph ? 2 * li + 1 : nfi), dj = (sph ? 2 * lj + 1 : nfj);
        const int dk = (sph ? 2 * lk + 1 : nfk), dl = (ncenter == 4) ? (sph ? 2 * ll + 1 : nfl) : 1;
        const size_t ni = dims ? dims[0] : (size_t)di * nci;
        const size_t nj = (ncenter == 2) ? 1 : dims ? dims[1] : (size_t)dj * ncj;
        const size_t nk = (ncenter == 2) ? 1 : dims ? dims[2] : (size_t)dk * nck;

        double *gctr = calloc((size_t)nc * nf, sizeof(double));
        int nonempty = 0;

        /* ---- shell-pair quantities ---- */
        const double *ri = sh[0].r, *rj = sh[1].r, *rk = sh[2].r, *rl = sh[3].r;
        double rr_ij = 0, rr_kl = 0;
        for (n = 0; n < 3; n++) {
                rr_ij += (ri[n] - rj[n]) * (ri[n] - rj[n]);
                rr_kl += (rk[n] - rl[n]) * (rk[n] - rl[n]);
        }
        /* log_rr terms of CINTset_pairdata (src/optimizer.c:301-314) and CINT2e_loop_nopt :117-147 */
        double log_rr_ij = 1.7 - 1.5 * log(sh[0].a[sh[0].nprim - 1] + sh[1].a[sh[1].nprim - 1]);
        double log_rr_kl = 1.7 - 1.5 * log(sh[2].a[sh[2].nprim - 1] + sh[3].a[sh[3].nprim - 1]);
        double cut_ij_extra = 0;
        {
                int lij = li + lj, lkl = lk + ll;
                double dist_ij = sqrt(rr_ij), dist_kl = sqrt(rr_kl);
                if (omega < 0) {
                        double om2 = omega * omega, r_guess = 8.;
                        double th_ij = om2 / (om2 + sh[0].a[sh[0].nprim - 1] + sh[1].a[sh[1].nprim - 1]);
                        if (lij > 0) log_rr_ij += lij * log(dist_ij + th_ij * r_guess + 1.);
                        if (order > 1) {
                                if (lij > 0) cut_ij_extra = lij * log((dist_ij + th_ij * r_guess + 1.) / (dist_ij + 1.));
                                if (ncenter == 4) {
                                        double th_kl = om2 / (om2 + sh[2].a[sh[2].nprim - 1] + sh[3].a[sh[3].nprim - 1]);
                                        if (lkl > 0) log_rr_kl += lkl * log(dist_kl + th_kl * r_guess + 1.);
                                } else if (lk > 0) {
                                        double th_k = om2 / (om2 + sh[2].a[sh[2].nprim - 1]);
                                        cut_ij_extra += lk * log(th_k * r_guess + 1.);
                                }
                        } else if (ncenter == 4 && lkl > 0) {
                                log_rr_kl += lkl * log(dist_kl + 1.);
                        }
                } else {
                        if (lij > 0) log_rr_ij += lij * log(dist_ij + 1.);
                        if (ncenter == 4 && lkl > 0) log_rr_kl += lkl * log(dist_kl + 1.);
                }
        }
        /* pair emptiness test uses the unmodified cutoff (CINTset_pairdata is called with envs->expcutoff) */
        {
                int any = 0, ip, jp;
                for (jp = 0; jp < sh[1].nprim; jp++)
                for (ip = 0; ip < sh[0].nprim; ip++) {
                        double ai = sh[0].a[ip], aj = sh[1].a[jp];
                        double cce = rr_ij * ai * aj / (ai + aj) - log_rr_ij - sh[0].logmaxc[ip] - sh[1].logmaxc[jp];
                        if (cce < expcutoff) any = 1;
                }
                if (!any) goto finish;
        }
        const double cutoff_loop = expcutoff + cut_ij_extra;

        /* ---- g-array geometry (src/g2e.c:101-151) ---- */
        const int ibase = li > lj, kbase = lk > ll;
        const int nmax = li + lj, mmax = lk + ll;
        GDim gd;
        gd.nr = nroots;
        gd.di = ibase ? nmax + 1 : li + 1;
        gd.dj = ibase ? lj + 1 : nmax + 1;
        gd.dk = kbase ? mmax + 1 : lk + 1;
        gd.dl = kbase ? ll + 1 : mmax + 1;
        const size_t gsize = (size_t)gd.nr * gd.di * gd.dj * gd.dk * gd.dl;
        double *g = malloc(sizeof(double) * gsize * 3);
        double *gout = malloc(sizeof(double) * nf);
        const double *rbra = ibase ? ri : rj;           /* rx_in_rijrx */
        const double *rket = kbase ? rk : rl;           /* rx_in_rklrx */
        double rirj[3], rkrl[3];
        for (n = 0; n < 3; n++) {
                rirj[n] = ibase ? ri[n] - rj[n] : rj[n] - ri[n];
                rkrl[n] = kbase ? rk[n] - rl[n] : rl[n] - rk[n];
        }
        int inx[128], iny[128], inz[128], jnx[128], jny[128], jnz[128];
        int knx[128], kny[128], knz[128], lnx[128], lny[128], lnz[128];
        cart_comp(li, inx, iny, inz);
        cart_comp(lj, jnx, jny, jnz);
        cart_comp(lk, knx, kny, knz);
        cart_comp(ll, lnx, lny, lnz);

        int ip, jp, kp, lp;
        for (lp = 0; lp < sh[3].nprim; lp++)
        for (kp = 0; kp < sh[2].nprim; kp++) {
                const double ak = sh[2].a[kp], al = sh[3].a[lp];
                const double akl = ak + al;
                double ekl_exp = rr_kl * ak * al / akl;
                double ccekl = 0;
                if (ncenter == 4) {
                        ccekl = ekl_exp - log_rr_kl - sh[2].logmaxc[kp] - sh[3].logmaxc[lp];
                        if (ccekl > cutoff_loop) continue;
                }
                double rkl[3];
                for (n = 0; n < 3; n++) rkl[n] = (ak * rk[n] + al * rl[n]) / akl;
                const double eijcutoff = cutoff_loop - ccekl;
                const double ekl = exp(-ekl_exp);
                for (jp = 0; jp < sh[1].nprim; jp++)
                for (ip = 0; ip < sh[0].nprim; ip++) {
                        const double ai = sh[0].a[ip], aj = sh[1].a[jp];
                        const double aij = ai + aj;
                        const double eij_exp = rr_ij * ai * aj / aij;
                        const double cceij = eij_exp - log_rr_ij - sh[0].logmaxc[ip] - sh[1].logmaxc[jp];
                        if (cceij > eijcutoff) continue;
                        if (!(cceij < expcutoff)) continue;   /* dead pair entry (rij = 1e18, eij = 0) */
                        double rij[3];
                        for (n = 0; n < 3; n++) rij[n] = ri[n] + aj / aij * (rj[n] - ri[n]);
                        const double cutoff = eijcutoff - cceij;
                        const double fac = common * exp(-eij_exp) * ekl;

                        /* ---- CINTg0_2e ---- */
                        double u[MAXROOTS], w[MAXROOTS];
                        double dx[3], rr = 0;
                        for (n = 0; n < 3; n++) { dx[n] = rij[n] - rkl[n]; rr += dx[n] * dx[n]; }
                        const double a1 = aij * akl;
                        const double a0 = a1 / (aij + akl);
                        double fac1 = sqrt(a0 / (a1 * a1 * a1)) * fac;
                        double x = a0 * rr;
                        if (omega == 0) {
                                if (oracle_rys_roots(nroots, x, u, w)) goto fail;
                        } else if (omega < 0) {
                                double theta = omega * omega / (omega * omega + a0);
                                if (theta * x > cutoff || theta * x > 40) continue;
                                if (order == nroots) {
                                        if (oracle_sr_rys_roots(nroots, x, sqrt(theta), u, w)) goto fail;
                                } else {
                                        int ir;
                                        double sqrt_theta = -sqrt(theta);
                                        if (oracle_rys_roots(order, x, u, w)) goto fail;
                                        if (oracle_rys_roots(order, theta * x, u + order, w + order)) goto fail;
                                        for (ir = order; ir < nroots; ir++) {
                                                double ut = u[ir] * theta;
                                                u[ir] = ut / (u[ir] + 1. - ut);
                                                w[ir] *= sqrt_theta;
                                        }
                                }
                        } else {
                                int ir;
                                double theta = omega * omega / (omega * omega + a0);
                                x *= theta;
                                fac1 *= sqrt(theta);
                                if (oracle_rys_roots(nroots, x, u, w)) goto fail;
                                for (ir = 0; ir < nroots; ir++) {
                                        double ut = u[ir] * theta;
                                        u[ir] = ut / (u[ir] + 1. - ut);
                                }
                        }

                        int ir, xyz;
                        for (ir = 0; ir < nroots; ir++) {
                                const double u2 = a0 * u[ir];
                                const double tmp4 = .5 / (u2 * (aij + akl) + a1);
                                const double b00 = u2 * tmp4;
                                const double b10 = b00 + tmp4 * akl;
                                const double b01 = b00 + tmp4 * aij;
                                for (xyz = 0; xyz < 3; xyz++) {
                                        double *gx = g + gsize * xyz;
                                        const double c00 = (rij[xyz] - rbra[xyz]) - 2 * b00 * akl * dx[xyz];
                                        const double c0p = (rkl[xyz] - rket[xyz]) + 2 * b00 * aij * dx[xyz];
                                        /* 2-D VRR on the carrier indices: G[n][m], n<=nmax on bra carrier, m<=mmax on ket carrier */
                                        double G[2 * LMAX_ORACLE + 2][2 * LMAX_ORACLE + 2];
                                        int nn, mm;
                                        G[0][0] = (xyz == 2) ? w[ir] * fac1 : 1.0;
                                        if (nmax > 0) G[1][0] = c00 * G[0][0];
                                        for (nn = 1; nn < nmax; nn++) G[nn + 1][0] = c00 * G[nn][0] + nn * b10 * G[nn - 1][0];
                                        for (mm = 0; mm < mmax; mm++)
                                                for (nn = 0; nn <= nmax; nn++) {
                                                        double v = c0p * G[nn][mm];
                                                        if (mm > 0) v += mm * b01 * G[nn][mm - 1];
                                                        if (nn > 0) v += nn * b00 * G[nn - 1][mm];
                                                        G[nn][mm + 1] = v;
                                                }
                                        /* place on the carriers, then 4-D HRR */
                                        int i, j, k, l;
                                        for (mm = 0; mm <= mmax; mm++)
                                        for (nn = 0; nn <= nmax; nn++) {
                                                i = ibase ? nn : 0; j = ibase ? 0 : nn;
                                                k = kbase ? mm : 0; l = kbase ? 0 : mm;
                                                gx[gidx(&gd, ir, i, j, k, l)] = G[nn][mm];
                                        }
                                        /* bra: move from carrier to the other index, for every ket carrier value */
                                        for (mm = 0; mm <= mmax; mm++) {
                                                k = kbase ? mm : 0; l = kbase ? 0 : mm;
                                                if (ibase) {
                                                        for (j = 1; j <= lj; j++)
                                                        for (i = 0; i <= nmax - j; i++)
                                                                gx[gidx(&gd, ir, i, j, k, l)] =
                                                                        rirj[xyz] * gx[gidx(&gd, ir, i, j - 1, k, l)]
                                                                        + gx[gidx(&gd, ir, i + 1, j - 1, k, l)];
                                                } else {
                                                        for (i = 1; i <= li; i++)
                                                        for (j = 0; j <= nmax - i; j++)
                                                                gx[gidx(&gd, ir, i, j, k, l)] =
                                                                        rirj[xyz] * gx[gidx(&gd, ir, i - 1, j, k, l)]
                                                                        + gx[gidx(&gd, ir, i - 1, j + 1, k, l)];
                                                }
                                        }
                                        /* ket */
                                        for (j = 0; j <= lj; j++)
                                        for (i = 0; i <= li; i++) {
                                                if (kbase) {
                                                        for (l = 1; l <= ll; l++)
                                                        for (k = 0; k <= mmax - l; k++)
                                                                gx[gidx(&gd, ir, i, j, k, l)] =
                                                                        rkrl[xyz] * gx[gidx(&gd, ir, i, j, k, l - 1)]
                                                                        + gx[gidx(&gd, ir, i, j, k + 1, l - 1)];
                                                } else {
                                                        for (k = 1; k <= lk; k++)
                                                        for (l = 0; l <= mmax - k; l++)
                                                                gx[gidx(&gd, ir, i, j, k, l)] =
                                                                        rkrl[xyz] * gx[gidx(&gd, ir, i, j, k - 1, l)]
                                                                        + gx[gidx(&gd, ir, i, j, k - 1, l + 1)];
                                                }
                                        }
                                }
                        }
                        /* ---- gout: n = i + nfi*(k + nfk*(l + nfl*j)) (src/g2e.c:206-265) ---- */
                        {
                                int i, j, k, l;
                                const double *gx = g, *gy = g + gsize, *gz = g + 2 * gsize;
                                n = 0;
                                for (j = 0; j < nfj; j++)
                                for (l = 0; l < nfl; l++)
                                for (k = 0; k < nfk; k++)
                                for (i = 0; i < nfi; i++, n++) {
                                        double s = 0;
                                        for (ir = 0; ir < nroots; ir++)
                                                s += gx[gidx(&gd, ir, inx[i], jnx[j], knx[k], lnx[l])]
                                                   * gy[gidx(&gd, ir, iny[i], jny[j], kny[k], lny[l])]
                                                   * gz[gidx(&gd, ir, inz[i], jnz[j], knz[k], lnz[l])];
                                        gout[n] = s;
                                }
                        }
                        /* ---- contraction ---- */
                        {
                                int ci, cj, ck, cl;
                                for (cl = 0; cl < ncl; cl++)
                                for (ck = 0; ck < nck; ck++)
                                for (cj = 0; cj < ncj; cj++)
                                for (ci = 0; ci < nci; ci++) {
                                        double cc = sh[0].c[ci * sh[0].nprim + ip] * sh[1].c[cj * sh[1].nprim + jp]
                                                  * sh[2].c[ck * sh[2].nprim + kp] * sh[3].c[cl * sh[3].nprim + lp];
                                        if (cc == 0) continue;
                                        double *dst = gctr + (size_t)(((cl * nck + ck) * ncj + cj) * nci + ci) * nf;
                                        for (n = 0; n < nf; n++) dst[n] += cc * gout[n];
                                }
                        }
                        nonempty = 1;
                }
        }
        free(g);
        free(gout);

finish:
        /* ---- cart->sph and scatter: out[i + ni*(j + nj*(k + nk*l))], contraction index major ---- */
        {
                double *ci_m = malloc(sizeof(double) * 4 * 21 * 128);
                double *cm[4];
                int ls[4] = {li, lj, lk, ll};
                int dd[4] = {di, dj, dk, dl}, nfc[4] = {nfi, nfj, nfk, nfl};
                for (n = 0; n < 4; n++) {
                        cm[n] = ci_m + n * 21 * 128;
                        if (sph) oracle_c2s_matrix(ls[n], cm[n]);
                }
                double *t1 = malloc(sizeof(double) * (size_t)nf * 2 + 16);
                double *t2 = t1 + nf;
                int ci, cj, ck, cl;
                for (cl = 0; cl < ncl; cl++)
                for (ck = 0; ck < nck; ck++)
                for (cj = 0; cj < ncj; cj++)
                for (ci = 0; ci < nci; ci++) {
                        const double *src = gctr + (size_t)(((cl * nck + ck) * ncj + cj) * nci + ci) * nf;
                        /* cart block is indexed [j][l][k][i] (i fastest); transform one index at a time */
                        int cur[4] = {nfi, nfj, nfk, nfl};      /* current extents i,j,k,l */
                        memcpy(t1, src, sizeof(double) * nf);
                        int which;
                        for (which = 0; which < 4 && sph; which++) {
                                int nw = dd[which], ow = nfc[which];
                                int e[4] = {cur[0], cur[1], cur[2], cur[3]};
                                int i, j, k, l, mm, cc;
                                int ne[4] = {e[0], e[1], e[2], e[3]};
                                ne[which] = nw;
                                for (j = 0; j < ne[1]; j++) for (l = 0; l < ne[3]; l++)
                                for (k = 0; k < ne[2]; k++) for (i = 0; i < ne[0]; i++) {
                                        int id[4] = {i, j, k, l};
                                        mm = id[which];
                                        double s = 0;
                                        for (cc = 0; cc < ow; cc++) {
                                                int is[4] = {i, j, k, l};
                                                is[which] = cc;
                                                s += cm[which][mm * ow + cc]
                                                   * t1[is[0] + e[0] * (is[2] + e[2] * (is[3] + e[3] * is[1]))];
                                        }
                                        t2[i + ne[0] * (k + ne[2] * (l + ne[3] * j))] = s;
                                }
                                cur[which] = nw;
                                { double *tt = t1; t1 = t2; t2 = tt; }
                        }
                        int i, j, k, l;
                        for (l = 0; l < dl; l++) for (k = 0; k < dk; k++)
                        for (j = 0; j < dj; j++) for (i = 0; i < di; i++)
                                out[(ci * di + i) + ni * ((cj * dj + j) + nj * ((ck * dk + k) + nk * (size_t)(cl * dl + l)))]
                                        = t1[i + di * (k + dk * (l + dl * j))];
                }
                if (t1 > t2) t1 = t2;
                free(t1);
                free(ci_m);
        }
        free(gctr);
        return nonempty;
fail:
        free(g);
        free(gout);
        free(gctr);
        return -1;
}

int oracle_int2e_sph(double *out, const int *dims, const int *shls, const int *atm, int natm,
                     const int *bas, int nbas, const double *env)
{ return eri_tuple(out, dims, shls, 4, 1, atm, natm, bas, nbas, env); }
int oracle_int2e_cart(double *out, const int *dims, const int *shls, const int *atm, int natm,
                      const int *bas, int nbas, const double *env)
{ return eri_tuple(out, dims, shls, 4, 0, atm, natm, bas, nbas, env); }
int oracle_int3c2e_sph(double *out, const int *dims, const int *shls, const int *atm, int natm,
                       const int *bas, int nbas, const double *env)
{ return eri_tuple(out, dims, shls, 3, 1, atm, natm, bas, nbas, env); }
int oracle_int3c2e_cart(double *out, const int *dims, const int *shls, const int *atm, int natm,
                        const int *bas, int nbas, const double *env)
{ return eri_tuple(out, dims, shls, 3, 0, atm, natm, bas, nbas, env); }
int oracle_int2c2e_sph(double *out, const int *dims, const int *shls, const int *atm, int natm,
                       const int *bas, int nbas, const double *env)
{ return eri_tuple(out, dims, shls, 2, 1, atm, natm, bas, nbas, env); }
int oracle_int2c2e_cart(double *out, const int *dims, const int *shls, const int *atm, int natm,
                        const int *bas, int nbas, const double *env)
{ return eri_tuple(out, dims, shls, 2, 0, atm, natm, bas, nbas, env); }

/* ---------------------------------------------------------------------------------------------
 * First derivative ( nabla i j | k l ) -- int2e_ip1 / int3c2e_ip1 (src/autocode/grad2.c:19-68,
 * src/autocode/int3c2e.c; ng = {1,0,0,0,1,1,1,3}).  The reference applies, on its g array,
 *     f(i) = i g(i-1) - 2 a_i g(i+1)                         (CINTnabla1i_2e, src/g2e.c:4550)
 * i.e. d/dx of the Cartesian Gaussian of centre i.  Restated one level up: the same identity holds for whole
 * shells, so the derivative block is a combination of the Cartesian blocks of a raised shell (l+1, coefficients
 * -2 a_p c_kp) and a lowered shell (l-1, coefficients c_kp), evaluated by eri_tuple above; s and p functions carry
 * fac_sp(l) instead of a transformation (src/g1e.c:565-572), which is undone / re-applied for the changed l.
 * Output: out[comp][l][k][j][i], comp = x, y, z (the reference's tensor-component-major layout).
 */
static int cidx(int lx, int lz, int l) { int r = l - lx; return r * (r + 1) / 2 + lz; }

static void c2s_pass(const double *in, double *out, size_t pre, int nin, int nout, size_t post, const double *cm)
{
        size_t p, q;
        int m, c;
        for (p = 0; p < pre; p++)
        for (m = 0; m < nout; m++)
        for (q = 0; q < post; q++) {
                double s = 0;
                for (c = 0; c < nin; c++) s += cm[m * nin + c] * in[(p * nin + c) * post + q];
                out[(p * nout + m) * post + q] = s;
        }
}

/* dpos: position of the differentiated shell in the tuple (0: ip1; last: int3c2e_ip2 / int2c2e_ip2, src/autocode/int3c2e.c:99,408) */
static int eri_ip1(double *out, const int *dims, const int *shls, int ncenter, int sph,
                   const int *atm, int natm, const int *bas, int nbas, const double *env, int dpos)
{
        const int ish = shls[dpos];
        const int li = bas[ish * BAS_SLOTS + ANG_OF], npi = bas[ish * BAS_SLOTS + NPRIM_OF], nci = bas[ish * BAS_SLOTS + NCTR_OF];
        /* private copy of the basis with the raised (index nbas) and lowered (nbas + 1) shell appended */
        int *xb = malloc(sizeof(int) * (nbas + 2) * BAS_SLOTS);
        memcpy(xb, bas, sizeof(int) * nbas * BAS_SLOTS);
        memcpy(xb + nbas * BAS_SLOTS, bas + ish * BAS_SLOTS, sizeof(int) * BAS_SLOTS);
        memcpy(xb + (nbas + 1) * BAS_SLOTS, bas + ish * BAS_SLOTS, sizeof(int) * BAS_SLOTS);
        int nenv = 20 /* PTR_ENV_START */, n, m;
        for (n = 0; n < natm; n++) if (atm[n * ATM_SLOTS + PTR_COORD] + 3 > nenv) nenv = atm[n * ATM_SLOTS + PTR_COORD] + 3;
        for (n = 0; n < nbas; n++) {
                int e1 = bas[n * BAS_SLOTS + PTR_EXP] + bas[n * BAS_SLOTS + NPRIM_OF];
                int e2 = bas[n * BAS_SLOTS + PTR_COEFF] + bas[n * BAS_SLOTS + NPRIM_OF] * bas[n * BAS_SLOTS + NCTR_OF];
                if (e1 > nenv) nenv = e1;
                if (e2 > nenv) nenv = e2;
        }
        double *xe = malloc(sizeof(double) * (nenv + npi * nci));
        memcpy(xe, env, sizeof(double) * nenv);
        const double *ai = env + bas[ish * BAS_SLOTS + PTR_EXP], *ci = env + bas[ish * BAS_SLOTS + PTR_COEFF];
        for (n = 0; n < nci; n++) for (m = 0; m < npi; m++) xe[nenv + n * npi + m] = -2 * ai[m] * ci[n * npi + m];
        xb[nbas * BAS_SLOTS + ANG_OF] = li + 1;
        xb[nbas * BAS_SLOTS + PTR_COEFF] = nenv;
        xb[(nbas + 1) * BAS_SLOTS + ANG_OF] = li > 0 ? li - 1 : 0;

        int dcart[4] = {1, 1, 1, 1}, nctr[4] = {1, 1, 1, 1}, ls[4] = {0, 0, 0, 0};
        for (n = 0; n < ncenter; n++) {
                ls[n] = bas[shls[n] * BAS_SLOTS + ANG_OF];
                nctr[n] = bas[shls[n] * BAS_SLOTS + NCTR_OF];
                dcart[n] = ncart(ls[n]) * nctr[n];
        }
        const int nfi = ncart(li), nfp = ncart(li + 1), nfm = li > 0 ? ncart(li - 1) : 0;
        size_t post = 1, rest = 1;
        for (n = 0; n < dpos; n++) post *= dcart[n];
        for (n = dpos + 1; n < 4; n++) rest *= dcart[n];
        double *bp = calloc((size_t)nfp * nci * rest * post, sizeof(double));
        double *bm = calloc((size_t)(nfm > 0 ? nfm : 1) * nci * rest * post, sizeof(double));
        int xs[4] = {shls[0], shls[1], ncenter > 2 ? shls[2] : 0, ncenter > 3 ? shls[3] : 0};
        xs[dpos] = nbas;
        int ret = eri_tuple(bp, NULL, xs, ncenter, 0, atm, natm, xb, nbas + 2, xe);
        if (li > 0) {
                xs[dpos] = nbas + 1;
                int r2 = eri_tuple(bm, NULL, xs, ncenter, 0, atm, natm, xb, nbas + 2, xe);
                if (r2 < 0) ret = r2; else if (ret >= 0) ret |= r2;
        }
        if (ret < 0) { free(xb); free(xe); free(bp); free(bm); return ret; }
        const double fi = fac_sp(li), sp = fi / fac_sp(li + 1), sm = li > 0 ? fi / fac_sp(li - 1) : 0;
        int cx[64], cy[64], cz[64];
        cart_comp(li, cx, cy, cz);
        /* Cartesian derivative blocks, then cart->sph index by index */
        const size_t ncar = (size_t)nfi * nci * rest * post;
        double *t1 = malloc(sizeof(double) * ncar * 2), *t2 = t1 + ncar;
        double *cmat = malloc(sizeof(double) * 21 * 128);
        int comp;
        for (comp = 0; comp < 3; comp++) {
                size_t r, q;
                int ic, a;
                for (r = 0; r < rest; r++) for (ic = 0; ic < nci; ic++) for (a = 0; a < nfi; a++) for (q = 0; q < post; q++) {
                        int nn = comp == 0 ? cx[a] : comp == 1 ? cy[a] : cz[a];
                        int up = cidx(cx[a] + (comp == 0), cz[a] + (comp == 2), li + 1);
                        double v = sp * bp[q + post * ((size_t)ic * nfp + up + (size_t)nci * nfp * r)];
                        if (nn > 0) v += nn * sm * bm[q + post * ((size_t)ic * nfm + cidx(cx[a] - (comp == 0), cz[a] - (comp == 2), li - 1) + (size_t)nci * nfm * r)];
                        t1[q + post * ((size_t)ic * nfi + a + (size_t)nci * nfi * r)] = v;
                }
                int d[4];
                for (n = 0; n < 4; n++) d[n] = dcart[n];
                if (sph) {
                        for (n = 0; n < ncenter; n++) {
                                if (ls[n] < 2) continue;
                                size_t post = 1, pre = nctr[n];
                                for (m = 0; m < n; m++) post *= d[m];
                                for (m = n + 1; m < 4; m++) pre *= d[m];
                                oracle_c2s_matrix(ls[n], cmat);
                                c2s_pass(t1, t2, pre, ncart(ls[n]), 2 * ls[n] + 1, post, cmat);
                                d[n] = (2 * ls[n] + 1) * nctr[n];
                                { double *tt = t1; t1 = t2; t2 = tt; }
                        }
                }
                const size_t ni = dims ? (size_t)dims[0] : (size_t)d[0], nj = dims ? (size_t)dims[1] : (size_t)d[1];
                const size_t nk = (dims && ncenter > 2) ? (size_t)dims[2] : (size_t)d[2], nl = (dims && ncenter > 3) ? (size_t)dims[3] : (size_t)d[3];
                /* 2-centre tuples are (i|k): the second shell sits at position 1 of d[] */
                int i, j, k, l;
                for (l = 0; l < d[3]; l++) for (k = 0; k < d[2]; k++) for (j = 0; j < d[1]; j++) for (i = 0; i < d[0]; i++)
                        out[comp * ni * nj * nk * nl + i + ni * (j + nj * (k + nk * (size_t)l))] = t1[i + (size_t)d[0] * (j + (size_t)d[1] * (k + (size_t)d[2] * l))];
        }
        if (t1 > t2) t1 = t2;
        free(t1); free(cmat); free(xb); free(xe); free(bp); free(bm);
        return ret;
}

int oracle_int2e_ip1_sph(double *out, const int *dims, const int *shls, const int *atm, int natm,
                         const int *bas, int nbas, const double *env)
{ return eri_ip1(out, dims, shls, 4, 1, atm, natm, bas, nbas, env, 0); }
int oracle_int2e_ip1_cart(double *out, const int *dims, const int *shls, const int *atm, int natm,
                          const int *bas, int nbas, const double *env)
{ return eri_ip1(out, dims, shls, 4, 0, atm, natm, bas, nbas, env, 0); }
int oracle_int3c2e_ip1_sph(double *out, const int *dims, const int *shls, const int *atm, int natm,
                           const int *bas, int nbas, const double *env)
{ return eri_ip1(out, dims, shls, 3, 1, atm, natm, bas, nbas, env, 0); }
int oracle_int3c2e_ip1_cart(double *out, const int *dims, const int *shls, const int *atm, int natm,
                            const int *bas, int nbas, const double *env)
{ return eri_ip1(out, dims, shls, 3, 0, atm, natm, bas, nbas, env, 0); }
int oracle_int3c2e_ip2_sph(double *out, const int *dims, const int *shls, const int *atm, int natm,
                           const int *bas, int nbas, const double *env)
{ return eri_ip1(out, dims, shls, 3, 1, atm, natm, bas, nbas, env, 2); }
int oracle_int2c2e_ip1_sph(double *out, const int *dims, const int *shls, const int *atm, int natm,
                           const int *bas, int nbas, const double *env)
{ return eri_ip1(out, dims, shls, 2, 1, atm, natm, bas, nbas, env, 0); }
int oracle_int2c2e_ip2_sph(double *out, const int *dims, const int *shls, const int *atm, int natm,
                           const int *bas, int nbas, const double *env)
{ return eri_ip1(out, dims, shls, 2, 1, atm, natm, bas, nbas, env, 1); }
