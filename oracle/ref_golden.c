/*
 * TEST INFRASTRUCTURE (never linked into or called by the product): whole-job fingerprints of the
 * UNMODIFIED reference library for the loop of examples/time_c60.c:200-219 (i >= j, all pairs kl with k <= i),
 * so that EVERY block of a whole-job run on the GPU can be value-checked, not just sampled ones.
 *
 * For every bra shell pair p = i(i+1)/2 + j, summed over all kets (k <= i, l <= k) and all block elements
 * v[mi,mj,mk,ml] (reference buf order, i fastest; dims include the contraction index):
 *     S[p] = sum v            A[p] = sum |v|
 *     F[p] = sum v * h(mi + di*mj) * g(ao_k + mk, ao_l + ml),   h(r) = cos(0.91 r + 0.3),  g(c,d) = cos(0.37 c + 0.61 d + 0.5)
 * (ao_* = spherical AO offset of the shell).  S/A/F are independent of how a GPU job chunks its tiles or shards its
 * kets over ranks: partial sums over the ranks' kets add up to them.
 *
 * Coulomb and exchange matrices for the fixed symmetric density D[a,b] = cos(0.37 (a+b) + 0.2) + 0.5 cos(0.11 (a-b)),
 *     J[a,b] = sum_cd (ab|cd) D[c,d]      K[a,c] = sum_bd (ab|cd) D[b,d]
 * digested from the 8-fold unique quartets (ij >= kl in pair order; the loop's redundant quartets k == i, l > j are
 * skipped), all eight index images written out explicitly -- the plain statement the device digestion is checked against.
 *
 * usage: ref_golden <lib.so> <basis.bin> <out.bin> [stride phase]
 * out.bin: int64 npair, nao; double S[npair], A[npair], F[npair], J[nao*nao], K[nao*nao] (row-major)
 * With stride > 1 only every stride-th bra pair is evaluated (S/A/F of the others stay 0, J/K are then partial sums --
 * used by the small-molecule tests only with stride 1).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <dlfcn.h>
#include <omp.h>

typedef int (*intor_t)(double *, int *, int *, int *, int, int *, int, double *, void *, double *);
typedef void (*optim_t)(void **, int *, int, int *, int, double *);

static double hfun(int r) { return cos(0.91 * r + 0.3); }
static double gfun(int c, int d) { return cos(0.37 * c + 0.61 * d + 0.5); }
static double dfun(int a, int b) { return cos(0.37 * (a + b) + 0.2) + 0.5 * cos(0.11 * (a - b)); }

int main(int argc, char **argv)
{
        if (argc < 4) { fprintf(stderr, "usage: %s lib.so basis.bin out.bin [stride phase]\n", argv[0]); return 2; }
        void *h = dlopen(argv[1], RTLD_NOW);
        if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 1; }
        intor_t intor = (intor_t)dlsym(h, "int2e_sph");
        optim_t optim = (optim_t)dlsym(h, "int2e_optimizer");
        FILE *f = fopen(argv[2], "rb");
        if (!f || !intor || !optim) { fprintf(stderr, "bad input\n"); return 1; }
        int hdr[3];
        if (fread(hdr, sizeof(int), 3, f) != 3) return 1;
        int natm = hdr[0], nbas = hdr[1], nenv = hdr[2];
        int *atm = malloc(sizeof(int) * natm * 6), *bas = malloc(sizeof(int) * nbas * 8);
        double *env = malloc(sizeof(double) * nenv);
        if (fread(atm, sizeof(int), natm * 6, f) != (size_t)natm * 6) return 1;
        if (fread(bas, sizeof(int), nbas * 8, f) != (size_t)nbas * 8) return 1;
        if (fread(env, sizeof(double), nenv, f) != (size_t)nenv) return 1;
        fclose(f);
        long stride = argc > 5 ? atol(argv[4]) : 1, phase = argc > 5 ? atol(argv[5]) : 0;

        long npair = (long)nbas * (nbas + 1) / 2;
        int *ish = malloc(sizeof(int) * npair), *jsh = malloc(sizeof(int) * npair);
        int *dim = malloc(sizeof(int) * nbas), *ao = malloc(sizeof(int) * (nbas + 1));
        long ij = 0;
        int maxd = 0;
        ao[0] = 0;
        for (int i = 0; i < nbas; i++) {
                dim[i] = (2 * bas[i * 8 + 1] + 1) * bas[i * 8 + 3];
                ao[i + 1] = ao[i] + dim[i];
                if (dim[i] > maxd) maxd = dim[i];
                for (int j = 0; j <= i; j++, ij++) { ish[ij] = i; jsh[ij] = j; }
        }
        const int nao = ao[nbas];
        double *D = malloc(sizeof(double) * nao * nao);
        for (int a = 0; a < nao; a++) for (int b = 0; b < nao; b++) D[(size_t)a * nao + b] = dfun(a, b);
        double *S = calloc(npair, sizeof(double)), *A = calloc(npair, sizeof(double)), *F = calloc(npair, sizeof(double));
        double *J = calloc((size_t)nao * nao, sizeof(double)), *K = calloc((size_t)nao * nao, sizeof(double));
        void *opt = NULL;
        optim(&opt, atm, natm, bas, nbas, env);
        double t0 = omp_get_wtime();
#pragma omp parallel
        {
                double *buf = malloc(sizeof(double) * maxd * maxd * maxd * maxd);
                double *Jt = calloc((size_t)nao * nao, sizeof(double)), *Kt = calloc((size_t)nao * nao, sizeof(double));
                double *hw = malloc(sizeof(double) * maxd * maxd);
                for (int r = 0; r < maxd * maxd; r++) hw[r] = hfun(r);
#pragma omp for schedule(dynamic, 2)
                for (long p = phase; p < npair; p += stride) {
                        const int i = ish[p], j = jsh[p], di = dim[i], dj = dim[j];
                        const long klmax = (long)(i + 1) * (i + 2) / 2;
                        double s_ = 0, a_ = 0, f_ = 0;
                        for (long kl = 0; kl < klmax; kl++) {
                                const int k = ish[kl], l = jsh[kl], dk = dim[k], dl = dim[l];
                                int shls[4] = {i, j, k, l};
                                intor(buf, NULL, shls, atm, natm, bas, nbas, env, opt, NULL);
                                const double sym = (kl > p) ? 0.0 : (i == j ? 0.5 : 1.0) * (k == l ? 0.5 : 1.0) * (kl == p ? 0.5 : 1.0);
                                for (int ml = 0; ml < dl; ml++)
                                for (int mk = 0; mk < dk; mk++) {
                                        const double gw = gfun(ao[k] + mk, ao[l] + ml);
                                        const int c = ao[k] + mk, d = ao[l] + ml;
                                        const double *v = buf + (size_t)di * dj * (mk + (size_t)dk * ml);
                                        double fs = 0;
                                        for (int r = 0; r < di * dj; r++) { s_ += v[r]; a_ += fabs(v[r]); fs += v[r] * hw[r]; }
                                        f_ += fs * gw;
                                        if (sym == 0.0) continue;
                                        for (int mj = 0; mj < dj; mj++)
                                        for (int mi = 0; mi < di; mi++) {
                                                const int a = ao[i] + mi, b = ao[j] + mj;
                                                const double x = sym * v[mi + di * mj];
#define DM(p_, q_) D[(size_t)(p_) * nao + (q_)]
#define JT(p_, q_) Jt[(size_t)(p_) * nao + (q_)]
#define KT(p_, q_) Kt[(size_t)(p_) * nao + (q_)]
                                                /* the eight index images (ab|cd) (ba|cd) (ab|dc) (ba|dc) (cd|ab) (dc|ab) (cd|ba) (dc|ba) */
                                                JT(a, b) += x * DM(c, d); JT(b, a) += x * DM(c, d); JT(a, b) += x * DM(d, c); JT(b, a) += x * DM(d, c);
                                                JT(c, d) += x * DM(a, b); JT(d, c) += x * DM(a, b); JT(c, d) += x * DM(b, a); JT(d, c) += x * DM(b, a);
                                                KT(a, c) += x * DM(b, d); KT(b, c) += x * DM(a, d); KT(a, d) += x * DM(b, c); KT(b, d) += x * DM(a, c);
                                                KT(c, a) += x * DM(d, b); KT(c, b) += x * DM(d, a); KT(d, a) += x * DM(c, b); KT(d, b) += x * DM(c, a);
                                        }
                                }
                        }
                        S[p] = s_; A[p] = a_; F[p] = f_;
                }
#pragma omp critical
                for (size_t n = 0; n < (size_t)nao * nao; n++) { J[n] += Jt[n]; K[n] += Kt[n]; }
                free(buf); free(Jt); free(Kt); free(hw);
        }
        double t1 = omp_get_wtime();
        FILE *o = fopen(argv[3], "wb");
        if (!o) { fprintf(stderr, "cannot write %s\n", argv[3]); return 1; }
        long long head[2] = {npair, nao};
        fwrite(head, sizeof(long long), 2, o);
        fwrite(S, sizeof(double), npair, o); fwrite(A, sizeof(double), npair, o); fwrite(F, sizeof(double), npair, o);
        fwrite(J, sizeof(double), (size_t)nao * nao, o); fwrite(K, sizeof(double), (size_t)nao * nao, o);
        fclose(o);
        double tot = 0;
        for (long p = 0; p < npair; p++) tot += S[p];
        printf("{\"seconds\": %.3f, \"npair\": %ld, \"nao\": %d, \"threads\": %d, \"sum\": %.15e}\n", t1 - t0, npair, nao, omp_get_max_threads(), tot);
        return 0;
}
