/*
 * Offline generator for the device Rys-root tables (libcint_b200/csrc/rys_tables.inc).
 *
 * What the product evaluates at run time (csrc/rys.cuh): for nroots = n and argument x
 *   x <  XMAX(n):  piecewise polynomials of degree DEG on a floating-point-exponent grid: with
 *                  v = x + C0 (C0 a power of two), every octave [C0 2^o, C0 2^(o+1)) of v is cut into
 *                  2^M equal intervals, so the interval index is just the top bits of v's IEEE
 *                  representation (no log, no division) and intervals widen as the functions get
 *                  smoother:  width = C0 2^o / 2^M.
 *                  t_k^2(x), w_k(x) = sum_j C[interval][j][2k|2k+1] * y^j,  y in [-1,1) inside the interval
 *   x >= XMAX(n):  asymptotic Gauss-Hermite form  t_k^2 = r_k / x,  w_k = v_k / sqrt(x)
 *                  (same switch point as the reference: x >= 35 + 5 n, src/rys_roots.c:67-78)
 * The reference instead uses hand-fitted piecewise formulas for n <= 5 and iterative solvers for
 * n >= 6; this table scheme is our own (north-star item 1: "per-thread FP64 root/weight evaluation
 * with coefficient tables staged in shared memory").
 *
 * Samples come from oracle/rys_stieltjes.c compiled with -DRYS_QUAD (__float128), the polynomial is
 * the Chebyshev interpolant converted to the monomial basis in quad precision and rounded once.
 * Every interval is verified on probe points against the quad solver using the same double Horner
 * arithmetic as the device; the worst errors are printed and written into the table header.
 *
 * build:  gcc -O2 -fopenmp -DRYS_QUAD tools/gen_rys_tables.c oracle/rys_stieltjes.c -Ioracle -lquadmath -lm
 * usage:  gen_rys_tables explore n deg C0 M  |  gen_rys_tables emit NMAX deg C0 M > rys_tables.inc
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <quadmath.h>

int rys_quad_t2w(int n, __float128 x, __float128 *t2, __float128 *w);

#define MAXDEG 15
#define MAXN 16
typedef __float128 Q;

static double xmax_of(int n) { return 35.0 + 5.0 * n; }

/* exponent grid: interval `it` covers x in [lo, lo + width) */
static double G_C0 = 8.0;
static int G_M = 4;
static void grid_interval(int it, double *lo, double *width)
{
        int o = it >> G_M, sub = it & ((1 << G_M) - 1);
        double base = ldexp(G_C0, o);
        *width = base / (1 << G_M);
        *lo = base + sub * (*width) - G_C0;
}
static int grid_count(int n)
{
        int it = 0;
        double lo, w;
        for (;; it++) {
                grid_interval(it, &lo, &w);
                if (lo >= xmax_of(n)) return it;
        }
}

/* fit one interval [a, a+h): coef[j][p], p = 2k (t2_k) or 2k+1 (w_k), monomials in y in [-1,1] */
static void fit_interval(int n, int deg, double a, double h, double *coef /* [(deg+1)][2n] */)
{
        int m = deg + 1, i, j, k, p;
        Q f[MAXDEG + 1][2 * MAXN];
        Q cheb[MAXDEG + 1][2 * MAXN];
        for (i = 0; i < m; i++) {
                Q y = cosq(M_PIq * (i + 0.5Q) / m);
                Q x = (Q)a + ((Q)h) * (y + 1) / 2;
                Q s[MAXN], w[MAXN];
                if (rys_quad_t2w(n, x, s, w)) { fprintf(stderr, "root failure n=%d\n", n); exit(1); }
                for (k = 0; k < n; k++) { f[i][2 * k] = s[k]; f[i][2 * k + 1] = w[k]; }
        }
        for (j = 0; j < m; j++)
                for (p = 0; p < 2 * n; p++) {
                        Q acc = 0;
                        for (i = 0; i < m; i++) acc += f[i][p] * cosq(M_PIq * j * (i + 0.5Q) / m);
                        cheb[j][p] = acc * (j == 0 ? 1.0Q : 2.0Q) / m;
                }
        /* Chebyshev -> monomial: T_0 = 1, T_1 = y, T_{j+1} = 2 y T_j - T_{j-1} */
        Q T[MAXDEG + 1][MAXDEG + 1];
        memset(T, 0, sizeof T);
        T[0][0] = 1;
        if (m > 1) T[1][1] = 1;
        for (j = 1; j + 1 < m; j++)
                for (k = 0; k <= j + 1; k++)
                        T[j + 1][k] = (k > 0 ? 2 * T[j][k - 1] : 0) - T[j - 1][k];
        for (p = 0; p < 2 * n; p++)
                for (k = 0; k < m; k++) {
                        Q acc = 0;
                        for (j = k; j < m; j++) acc += cheb[j][p] * T[j][k];
                        coef[k * 2 * n + p] = (double)acc;
                }
}

static void check_interval(int n, int deg, double a, double h, const double *coef, double *err_s, double *err_w)
{
        static const double probes[] = {-0.999, -0.93, -0.71, -0.33, 0.05, 0.41, 0.77, 0.95, 0.9999};
        int ip, k, j;
        for (ip = 0; ip < (int)(sizeof probes / sizeof probes[0]); ip++) {
                double y = probes[ip];
                Q x = (Q)a + ((Q)h) * ((Q)y + 1) / 2;
                Q s[MAXN], w[MAXN];
                rys_quad_t2w(n, x, s, w);
                for (k = 0; k < 2 * n; k++) {
                        double v = coef[deg * 2 * n + k];
                        for (j = deg - 1; j >= 0; j--) v = fma(v, y, coef[j * 2 * n + k]);
                        if (k % 2 == 0) {
                                double e = fabs((double)((Q)v - s[k / 2]));
                                if (e > *err_s) *err_s = e;
                        } else {
                                /* weight error measured against the total weight F_0(x) = sum_k w_k:
                                 * that is the scale on which it enters an integral */
                                Q tot = 0;
                                for (j = 0; j < n; j++) tot += w[j];
                                double e = fabs((double)(((Q)v - w[k / 2]) / tot));
                                if (e > *err_w) *err_w = e;
                        }
                }
        }
}

/* large-x constants: Gauss-Hermite.  For x -> inf the measure exp(-x t^2) on [0,1] becomes
 * exp(-z^2) dz/sqrt(x) on [0,inf) with z = sqrt(x) t, so t_k^2 = z_k^2 / x, w_k = v_k / sqrt(x)
 * where (z_k^2, v_k) is the n-point rule of exp(-z^2) on [0,inf) in z^2.  Obtained numerically from
 * the same solver: at X = 4000 the tail beyond t = 1 is < exp(-4000). */
static void largex(int n, double *r, double *v)
{
        Q s[MAXN], w[MAXN], X = 4000;
        int k;
        rys_quad_t2w(n, X, s, w);
        for (k = 0; k < n; k++) { r[k] = (double)(s[k] * X); v[k] = (double)(w[k] * sqrtq(X)); }
}

static double check_largex(int n, const double *r, const double *v, double *err_w)
{
        double xs[] = {0, 0.5, 3, 10, 40, 200, 1000};
        double es = 0;
        int i, k;
        for (i = 0; i < 7; i++) {
                double x = xmax_of(n) + xs[i];
                Q s[MAXN], w[MAXN];
                rys_quad_t2w(n, x, s, w);
                for (k = 0; k < n; k++) {
                        double e = fabs((double)((Q)(r[k] / x) - s[k]));
                        if (e > es) es = e;
                        Q tot = 0;
                        for (int j = 0; j < n; j++) tot += w[j];
                        e = fabs((double)(((Q)(v[k] / sqrt(x)) - w[k]) / tot));
                        if (e > *err_w) *err_w = e;
                }
        }
        return es;
}

int main(int argc, char **argv)
{
        if (argc >= 6 && !strcmp(argv[1], "explore")) {
                int n = atoi(argv[2]), deg = atoi(argv[3]);
                G_C0 = atof(argv[4]);
                G_M = atoi(argv[5]);
                int nint = grid_count(n), it;
                double es = 0, ew = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(max : es, ew)
                for (it = 0; it < nint; it++) {
                        double coef[(MAXDEG + 1) * 2 * MAXN], e1 = 0, e2 = 0, lo, h;
                        grid_interval(it, &lo, &h);
                        fit_interval(n, deg, lo, h, coef);
                        check_interval(n, deg, lo, h, coef, &e1, &e2);
                        if (e1 > es) es = e1;
                        if (e2 > ew) ew = e2;
                }
                printf("n=%d deg=%d C0=%g M=%d intervals=%d bytes=%d  max abs err t2 %.2e  max err w/F0 %.2e\n",
                       n, deg, G_C0, G_M, nint, nint * (deg + 1) * 2 * n * 8, es, ew);
                return 0;
        }
        if (argc >= 6 && !strcmp(argv[1], "emit")) {
                int nmax = atoi(argv[2]), deg = atoi(argv[3]), n;
                G_C0 = atof(argv[4]);
                G_M = atoi(argv[5]);
                printf("// GENERATED by tools/gen_rys_tables.c (emit %d %d %g %d) -- do not edit.\n", nmax, deg, G_C0, G_M);
                printf("// layout per nroots n: coef[interval][j = 0..DEG][p = 0..2n), p = 2k: t_k^2, p = 2k+1: w_k;\n");
                printf("// monomials in y in [-1,1) on the exponent grid of v = x + C0.  x >= 35+5n: t_k^2 = LX_R/x, w_k = LX_V/sqrt(x).\n");
                printf("#define RYS_TAB_NMAX %d\n#define RYS_TAB_DEG %d\n#define RYS_TAB_C0 %.17g\n#define RYS_TAB_M %d\n", nmax, deg, G_C0, G_M);
                long off = 0;
                long offs[MAXN + 2];
                int nints[MAXN + 2];
                double worst_s = 0, worst_w = 0;
                printf("static const double RYS_TAB_COEF[] = {\n");
                for (n = 1; n <= nmax; n++) {
                        int nint = grid_count(n), it;
                        size_t per = (size_t)(deg + 1) * 2 * n;
                        double *all = malloc(sizeof(double) * per * nint);
                        double es = 0, ew = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(max : es, ew)
                        for (it = 0; it < nint; it++) {
                                double e1 = 0, e2 = 0, lo, h;
                                grid_interval(it, &lo, &h);
                                fit_interval(n, deg, lo, h, all + per * it);
                                check_interval(n, deg, lo, h, all + per * it, &e1, &e2);
                                if (e1 > es) es = e1;
                                if (e2 > ew) ew = e2;
                        }
                        printf("// nroots %d: %d intervals, offset %ld, verified max abs err t^2 %.2e, max err w/F0 %.2e\n",
                               n, nint, off, es, ew);
                        fprintf(stderr, "n=%d intervals=%d err_t2=%.2e err_w=%.2e\n", n, nint, es, ew);
                        for (size_t i = 0; i < per * nint; i++)
                                printf("%.17g,%s", all[i], (i % 4 == 3) ? "\n" : " ");
                        printf("\n");
                        offs[n] = off;
                        nints[n] = nint;
                        off += per * nint;
                        if (es > worst_s) worst_s = es;
                        if (ew > worst_w) worst_w = ew;
                        free(all);
                }
                printf("};\nstatic const int RYS_TAB_OFF[RYS_TAB_NMAX + 1] = {0");
                for (n = 1; n <= nmax; n++) printf(", %ld", offs[n]);
                printf("};\nstatic const int RYS_TAB_NINT[RYS_TAB_NMAX + 1] = {0");
                for (n = 1; n <= nmax; n++) printf(", %d", nints[n]);
                printf("};\n");
                printf("// large-x constants, packed triangular: entry n(n-1)/2 + k\n");
                printf("static const double RYS_LX_R[] = {\n");
                double lw = 0, ls = 0;
                double rr[MAXN + 1][MAXN], vv[MAXN + 1][MAXN];
                for (n = 1; n <= nmax; n++) {
                        largex(n, rr[n], vv[n]);
                        double e = check_largex(n, rr[n], vv[n], &lw);
                        if (e > ls) ls = e;
                        for (int k = 0; k < n; k++) printf("%.17g, ", rr[n][k]);
                        printf("\n");
                }
                printf("};\nstatic const double RYS_LX_V[] = {\n");
                for (n = 1; n <= nmax; n++) {
                        for (int k = 0; k < n; k++) printf("%.17g, ", vv[n][k]);
                        printf("\n");
                }
                printf("};\n// verified: table max abs err t^2 %.2e, max err w/F0 %.2e; large-x %.2e / %.2e\n",
                       worst_s, worst_w, ls, lw);
                fprintf(stderr, "large-x err_t2=%.2e err_w=%.2e\n", ls, lw);
                return 0;
        }
        fprintf(stderr, "usage: %s explore n deg C0 M | emit nmax deg C0 M\n", argv[0]);
        return 2;
}
