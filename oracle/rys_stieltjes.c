/*
 * TEST INFRASTRUCTURE (CPU oracle) -- never linked into or called from the product path.
 *
 * Rys quadrature roots and weights from first principles, in extended precision.
 *
 * Definition followed (reference: src/rys_roots.c:57-122 `CINTrys_roots`, documented in
 * scripts/rys_roots.py:197 `rys_roots_weights`): for x >= 0 find nodes t_k^2 and weights w_k of the
 * n-point Gauss rule of the measure  exp(-x t^2) dt  on t in [0,1]  acting on polynomials in t^2,
 * and return them in the reference's convention   u_k = t_k^2 / (1 - t_k^2),  sum_k w_k = F_0(x).
 *
 * The reference reaches these numbers through piecewise polynomial fits (n <= 5), Wheeler/Schmidt
 * solvers and 80/128-bit fallbacks (n >= 6).  This restatement deliberately uses ONE independent,
 * textbook method for every n so it can cross-check both the reference and the GPU tables:
 *   1. discretise the measure with composite Gauss-Legendre panels in t (exponentially convergent
 *      for the entire integrand),
 *   2. Lanczos / RKPW (Gragg-Harrod, Gautschi ORTHPOL `lancz`) -> three-term recurrence (alpha, beta)
 *      of the polynomials orthogonal in s = t^2,
 *   3. Golub-Welsch: implicit-shift QL on the Jacobi matrix -> nodes s_k, weights beta_0 * v_k0^2.
 * Precision: `long double` by default (the oracle), `__float128` with -DRYS_QUAD (table generator).
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "eri_oracle.h"

#ifdef RYS_QUAD
#include <quadmath.h>
typedef __float128 REAL;
#define R_EXP  expq
#define R_SQRT sqrtq
#define R_FABS fabsq
#define R_EPS  1e-33Q
#define R_COS  cosq
#define R_PI   3.14159265358979323846264338327950288419716939937510Q
#else
typedef long double REAL;
#define R_EXP  expl
#define R_SQRT sqrtl
#define R_FABS fabsl
#define R_EPS  1e-19L
#define R_COS  cosl
#define R_PI   3.14159265358979323846264338327950288419716939937510L
#endif

#define GL_ORDER 32          /* Gauss-Legendre points per panel */
#define MAX_PANELS 24
#define MAXN 32

static REAL gl_x[GL_ORDER], gl_w[GL_ORDER];
static int gl_ready = 0;

/* Gauss-Legendre nodes on [-1,1] by Newton iteration on P_n (done once). */
static void gl_init(void)
{
        const int n = GL_ORDER;
        int i, k, it;
        for (i = 0; i < (n + 1) / 2; i++) {
                REAL z = R_COS(R_PI * (i + 0.75) / (n + 0.5));
                REAL pp = 0;
                for (it = 0; it < 100; it++) {
                        REAL p1 = 1, p2 = 0;
                        for (k = 0; k < n; k++) {
                                REAL p3 = p2;
                                p2 = p1;
                                p1 = ((2 * k + 1) * z * p2 - k * p3) / (k + 1);
                        }
                        pp = n * (z * p1 - p2) / (z * z - 1);
                        REAL dz = p1 / pp;
                        z -= dz;
                        if (R_FABS(dz) < R_EPS) break;
                }
                /* one more evaluation of P'_n at the converged node */
                {
                        REAL p1 = 1, p2 = 0;
                        for (k = 0; k < n; k++) {
                                REAL p3 = p2;
                                p2 = p1;
                                p1 = ((2 * k + 1) * z * p2 - k * p3) / (k + 1);
                        }
                        pp = n * (z * p1 - p2) / (z * z - 1);
                }
                gl_x[i] = -z;
                gl_x[n - 1 - i] = z;
                gl_w[i] = gl_w[n - 1 - i] = 2 / ((1 - z * z) * pp * pp);
        }
        gl_ready = 1;
}

/* Golub-Welsch on the symmetric tridiagonal (d, e): eigenvalues -> d, first eigenvector components -> z. */
static int jacobi_ql(int n, REAL *d, REAL *e, REAL *z)
{
        int l, m, i, iter;
        for (i = 0; i < n; i++) z[i] = 0;
        z[0] = 1;
        e[n - 1] = 0;
        for (l = 0; l < n; l++) {
                iter = 0;
                for (;;) {
                        for (m = l; m < n - 1; m++) {
                                REAL dd = R_FABS(d[m]) + R_FABS(d[m + 1]);
                                if (R_FABS(e[m]) <= R_EPS * dd) break;
                        }
                        if (m == l) break;
                        if (++iter > 200) return 1;
                        REAL g = (d[l + 1] - d[l]) / (2 * e[l]);
                        REAL r = R_SQRT(g * g + 1);
                        g = d[m] - d[l] + e[l] / (g + (g >= 0 ? r : -r));
                        REAL s = 1, c = 1, p = 0;
                        for (i = m - 1; i >= l; i--) {
                                REAL f = s * e[i];
                                REAL b = c * e[i];
                                r = R_SQRT(f * f + g * g);
                                e[i + 1] = r;
                                if (r == 0) {
                                        d[i + 1] -= p;
                                        e[m] = 0;
                                        break;
                                }
                                s = f / r;
                                c = g / r;
                                g = d[i + 1] - p;
                                r = (d[i] - g) * s + 2 * c * b;
                                p = s * r;
                                d[i + 1] = g + p;
                                g = c * r - b;
                                f = z[i + 1];
                                z[i + 1] = s * z[i] + c * f;
                                z[i] = c * z[i] - s * f;
                        }
                        if (r == 0 && i >= l) continue;
                        d[l] -= p;
                        e[l] = g;
                        e[m] = 0;
                }
        }
        return 0;
}

/*
 * Core: nodes s_k = t_k^2 (ascending) and weights w_k of the n-point rule for exp(-x t^2) on [0,1].
 * `lower` in [0,1): the measure is restricted to t in [lower, 1] (short-range/erfc variant,
 * reference src/rys_roots.c:149 `CINTsr_rys_roots`); lower = 0 is the plain Coulomb case.
 * The weights carry the factor exp(-x*lower^2)-free definition: w = int_lower^1 exp(-x t^2) ... dt.
 */
static int rys_core(int n, REAL x, REAL lower, REAL *s, REAL *w)
{
        static REAL xs[GL_ORDER * MAX_PANELS], ws[GL_ORDER * MAX_PANELS];
#pragma omp threadprivate(xs, ws)
        REAL p0[MAXN + 1], p1[MAXN + 1];
        int npanel, ip, i, k, N = 0;
        if (n < 1 || n > MAXN) return 2;
        if (!gl_ready) {
#pragma omp critical(rys_gl_init)
                if (!gl_ready) gl_init();
        }
        /* The integrand decays like exp(-x t^2): put the panels on [lower, tmax] with tmax where
         * the weight has dropped below 1e-40 of its start value; the remainder is negligible. */
        REAL tmax = 1;
        if (x > 0) {
                REAL tcut = R_SQRT(lower * lower + 92 / x);
                if (tcut < tmax) tmax = tcut;
        }
        npanel = 6 + 2 * n;
        if (npanel > MAX_PANELS) npanel = MAX_PANELS;
        REAL h = (tmax - lower) / npanel;
        for (ip = 0; ip < npanel; ip++) {
                REAL a = lower + ip * h;
                for (i = 0; i < GL_ORDER; i++, N++) {
                        REAL t = a + (gl_x[i] + 1) * h / 2;
                        xs[N] = t * t;
                        ws[N] = gl_w[i] * h / 2 * R_EXP(-x * (t * t - lower * lower));
                }
        }
        /* RKPW Lanczos, truncated to the first n recurrence coefficients. */
        for (k = 0; k <= n; k++) { p0[k] = 0; p1[k] = 0; }
        for (k = 0; k < n && k < N; k++) p0[k] = xs[k];
        p1[0] = ws[0];
        for (i = 0; i < N - 1; i++) {
                REAL pi = ws[i + 1], gam = 1, sig = 0, t = 0, xlam = xs[i + 1];
                int kmax = (i + 1 < n - 1) ? i + 1 : n - 1;
                if (i + 1 < n) p0[i + 1] = xs[i + 1];   /* entry becomes live at this step */
                for (k = 0; k <= kmax; k++) {
                        REAL rho = p1[k] + pi;
                        REAL tmp = gam * rho;
                        REAL tsig = sig;
                        if (rho <= 0) { gam = 1; sig = 0; }
                        else { gam = p1[k] / rho; sig = pi / rho; }
                        REAL tk = sig * (p0[k] - xlam) - gam * t;
                        p0[k] -= (tk - t);
                        t = tk;
                        if (sig <= 0) pi = tsig * p1[k];
                        else pi = (t * t) / sig;
                        p1[k] = tmp;
                }
        }
        /* Jacobi matrix: diag alpha_k = p0[k], off-diag sqrt(beta_k) = sqrt(p1[k]), k >= 1. */
        REAL d[MAXN], e[MAXN], z[MAXN];
        for (k = 0; k < n; k++) d[k] = p0[k];
        for (k = 0; k < n - 1; k++) e[k] = R_SQRT(p1[k + 1]);
        if (jacobi_ql(n, d, e, z)) return 1;
        /* sort ascending (insertion; n is tiny) */
        for (i = 1; i < n; i++) {
                REAL dv = d[i], zv = z[i];
                for (k = i - 1; k >= 0 && d[k] > dv; k--) { d[k + 1] = d[k]; z[k + 1] = z[k]; }
                d[k + 1] = dv; z[k + 1] = zv;
        }
        REAL scale = R_EXP(-x * lower * lower);
        for (k = 0; k < n; k++) {
                s[k] = d[k];
                w[k] = p1[0] * z[k] * z[k] * scale;
        }
        return 0;
}

/* t_k^2 and weights in double (rounded from extended precision). */
int oracle_rys_t2w(int n, double x, double *t2, double *w)
{
        REAL s[MAXN], ww[MAXN];
        int k, err = rys_core(n, (REAL)x, 0, s, ww);
        if (err) return err;
        for (k = 0; k < n; k++) { t2[k] = (double)s[k]; w[k] = (double)ww[k]; }
        return 0;
}

/* Reference convention (src/rys_roots.c:57): u = t^2/(1-t^2). */
int oracle_rys_roots(int n, double x, double *u, double *w)
{
        REAL s[MAXN], ww[MAXN];
        int k, err = rys_core(n, (REAL)x, 0, s, ww);
        if (err) return err;
        for (k = 0; k < n; k++) { u[k] = (double)(s[k] / (1 - s[k])); w[k] = (double)ww[k]; }
        return 0;
}

/*
 * Short-range roots, reference convention (src/rys_roots.c:149 `CINTsr_rys_roots(n,x,lower,u,w)`):
 * Gauss rule of exp(-x t^2) on t in [lower,1]; u = t^2/(1-t^2).
 */
int oracle_sr_rys_roots(int n, double x, double lower, double *u, double *w)
{
        REAL s[MAXN], ww[MAXN];
        int k, err = rys_core(n, (REAL)x, (REAL)lower, s, ww);
        if (err) return err;
        for (k = 0; k < n; k++) { u[k] = (double)(s[k] / (1 - s[k])); w[k] = (double)ww[k]; }
        return 0;
}

#ifdef RYS_QUAD
/* Extended-precision outputs for the table generator (tools/gen_rys_tables.c). */
int rys_quad_t2w(int n, __float128 x, __float128 *t2, __float128 *w) { return rys_core(n, x, 0, t2, w); }
#endif
