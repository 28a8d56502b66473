/*
 * TEST/BENCH INFRASTRUCTURE: times the UNMODIFIED reference library (oracle/_ref/libcint_ref.so, or
 * any library exporting the libcint ABI) on the host cores, with the loop shape of the reference's own
 * benchmark driver (examples/time_c60.c:196-219: OpenMP `for ij` schedule(dynamic,2), kl <= pairs with
 * k <= i, optimizer on).  Differences from that driver, all stated in the JSON it prints:
 *   - the basis comes from a binary dump (natm, nbas, nenv, atm, bas, env) instead of inline C;
 *   - only every `stride`-th ij pair (offset `phase`) is evaluated so the run is a bounded sample;
 *   - one buffer per thread instead of malloc/free per quartet (the library call is what is timed).
 * usage: time_ref <lib.so> <basis.bin> <stride> <phase> [aux_shell0]
 *        time_ref <lib.so> <basis.bin> ip1 <stride>        gradient loop of examples/time_c2h6.c:798-835: int2e_ip1_sph for every
 *                 stride-th ordered pair (i,j) and all k >= l, optimizer on
 *        time_ref <lib.so> <basis.bin> sweep <list.txt>     class sweep (BASELINE config 4): every line "i j k l reps" of the list
 *                 is timed as `reps` calls of int2e_sph on that shell quartet spread over the OpenMP threads (optimizer on)
 * With aux_shell0 > 0 the density-fitting loop is timed instead: int3c2e_sph for every `stride`-th orbital shell pair
 * i >= j < aux_shell0 and ALL auxiliary shells k >= aux_shell0 (the loop shape of cintb200_int3c2e_sph_all).
 */
#include <stdio.h>
#include <stdlib.h>
#include <dlfcn.h>
#include <omp.h>

typedef int (*intor_t)(double *, int *, int *, int *, int, int *, int, double *, void *, double *);
typedef void (*optim_t)(void **, int *, int, int *, int, double *);
typedef void (*delopt_t)(void **);

int main(int argc, char **argv)
{
        if (argc < 5) { fprintf(stderr, "usage: %s lib.so basis.bin stride phase\n", argv[0]); return 2; }
        void *h = dlopen(argv[1], RTLD_NOW);
        if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 1; }
        intor_t intor = (intor_t)dlsym(h, "int2e_sph");
        optim_t optim = (optim_t)dlsym(h, "int2e_optimizer");
        delopt_t delopt = (delopt_t)dlsym(h, "CINTdel_optimizer");
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
        if (argv[3][0] == 'i') {          /* first-derivative loop */
                intor_t ip1 = (intor_t)dlsym(h, "int2e_ip1_sph");
                optim_t oip1 = (optim_t)dlsym(h, "int2e_ip1_optimizer");
                if (!ip1 || !oip1) { fprintf(stderr, "no int2e_ip1_sph in %s\n", argv[1]); return 1; }
                long strd = atol(argv[4]);
                int md = 0;
                for (int i = 0; i < nbas; i++) { int d = (2 * bas[i * 8 + 1] + 1) * bas[i * 8 + 3]; if (d > md) md = d; }
                void *opt1 = NULL;
                oip1(&opt1, atm, natm, bas, nbas, env);
                double nints = 0, chk = 0;
                long nq = 0;
                double t0 = omp_get_wtime();
#pragma omp parallel reduction(+ : nints, chk, nq)
                {
                        double *buf = malloc(sizeof(double) * 3 * (size_t)md * md * md * md);
#pragma omp for schedule(dynamic, 2)
                        for (long ij = 0; ij < (long)nbas * nbas; ij += strd) {
                                int i = (int)(ij / nbas), j = (int)(ij - (long)nbas * i);
                                long dij = (long)(2 * bas[i * 8 + 1] + 1) * bas[i * 8 + 3] * (2 * bas[j * 8 + 1] + 1) * bas[j * 8 + 3];
                                for (int k = 0; k < nbas; k++)
                                        for (int l = 0; l <= k; l++) {
                                                int shls[4] = {i, j, k, l};
                                                ip1(buf, NULL, shls, atm, natm, bas, nbas, env, opt1, NULL);
                                                long n = 3 * dij * (2 * bas[k * 8 + 1] + 1) * bas[k * 8 + 3] * (2 * bas[l * 8 + 1] + 1) * bas[l * 8 + 3];
                                                nints += n;
                                                chk += buf[0] + buf[n - 1];
                                                nq++;
                                        }
                        }
                        free(buf);
                }
                double t1 = omp_get_wtime();
                printf("{\"seconds\": %.6f, \"integrals\": %.0f, \"quartets\": %ld, \"threads\": %d, \"stride\": %ld, \"phase\": 0, \"checksum\": %.15e}\n",
                       t1 - t0, nints, nq, omp_get_max_threads(), strd, chk);
                return 0;
        }
        if (argv[3][0] == 's') {          /* class sweep */
                FILE *lf = fopen(argv[4], "r");
                if (!lf) { fprintf(stderr, "cannot read %s\n", argv[4]); return 1; }
                int md = 0;
                for (int i = 0; i < nbas; i++) { int d = (2 * bas[i * 8 + 1] + 1) * bas[i * 8 + 3]; if (d > md) md = d; }
                void *opts = NULL;
                optim(&opts, atm, natm, bas, nbas, env);
                int q[4];
                long reps;
                printf("{\"threads\": %d, \"rows\": [", omp_get_max_threads());
                int first = 1;
                while (fscanf(lf, "%d %d %d %d %ld", &q[0], &q[1], &q[2], &q[3], &reps) == 5) {
                        double chk = 0;
                        double t0 = omp_get_wtime();
#pragma omp parallel reduction(+ : chk)
                        {
                                double *buf = malloc(sizeof(double) * (size_t)md * md * md * md);
#pragma omp for schedule(static)
                                for (long r = 0; r < reps; r++) {
                                        int shls[4] = {q[0], q[1], q[2], q[3]};
                                        intor(buf, NULL, shls, atm, natm, bas, nbas, env, opts, NULL);
                                        chk += buf[0];
                                }
                                free(buf);
                        }
                        double t1 = omp_get_wtime();
                        long n = 1;
                        for (int m = 0; m < 4; m++) n *= (2 * bas[q[m] * 8 + 1] + 1) * bas[q[m] * 8 + 3];
                        printf("%s[%d,%d,%d,%d,%ld,%.6e,%ld,%.6e]", first ? "" : ",", q[0], q[1], q[2], q[3], reps, t1 - t0, n, chk);
                        first = 0;
                }
                printf("]}\n");
                return 0;
        }
        long stride = atol(argv[3]), phase = atol(argv[4]);
        int aux0 = (argc > 5) ? atoi(argv[5]) : 0;
        if (aux0 > 0) {
                intor_t intor3 = (intor_t)dlsym(h, "int3c2e_sph");
                optim_t optim3 = (optim_t)dlsym(h, "int3c2e_optimizer");
                if (!intor3 || !optim3 || aux0 >= nbas) { fprintf(stderr, "bad 3-centre input\n"); return 1; }
                long np3 = (long)aux0 * (aux0 + 1) / 2;
                int *is3 = malloc(sizeof(int) * np3), *js3 = malloc(sizeof(int) * np3);
                long q = 0;
                int md = 0;
                for (int i = 0; i < nbas; i++) { int d = (2 * bas[i * 8 + 1] + 1) * bas[i * 8 + 3]; if (d > md) md = d; }
                for (int i = 0; i < aux0; i++) for (int j = 0; j <= i; j++, q++) { is3[q] = i; js3[q] = j; }
                void *opt3 = NULL;
                optim3(&opt3, atm, natm, bas, nbas, env);
                double nints3 = 0, chk3 = 0;
                long ntrip = 0;
                double t0 = omp_get_wtime();
#pragma omp parallel reduction(+ : nints3, chk3, ntrip)
                {
                        double *buf = malloc(sizeof(double) * md * md * md);
#pragma omp for schedule(dynamic, 2)
                        for (long p = phase; p < np3; p += stride) {
                                int i = is3[p], j = js3[p];
                                long dij = (long)(2 * bas[i * 8 + 1] + 1) * bas[i * 8 + 3] * (2 * bas[j * 8 + 1] + 1) * bas[j * 8 + 3];
                                for (int k = aux0; k < nbas; k++) {
                                        int shls[3] = {i, j, k};
                                        intor3(buf, NULL, shls, atm, natm, bas, nbas, env, opt3, NULL);
                                        long n = dij * (2 * bas[k * 8 + 1] + 1) * bas[k * 8 + 3];
                                        nints3 += n;
                                        chk3 += buf[0] + buf[n - 1];
                                }
                                ntrip += nbas - aux0;
                        }
                        free(buf);
                }
                double t1 = omp_get_wtime();
                printf("{\"seconds\": %.6f, \"integrals\": %.0f, \"quartets\": %ld, \"threads\": %d, \"stride\": %ld, \"phase\": %ld, \"checksum\": %.15e}\n",
                       t1 - t0, nints3, ntrip, omp_get_max_threads(), stride, phase, chk3);
                return 0;
        }
        long npair = (long)nbas * (nbas + 1) / 2;
        int *ish = malloc(sizeof(int) * npair), *jsh = malloc(sizeof(int) * npair);
        int *dim = malloc(sizeof(int) * nbas);
        long ij = 0;
        int maxd = 0;
        for (int i = 0; i < nbas; i++) {
                dim[i] = (2 * bas[i * 8 + 1] + 1) * bas[i * 8 + 3];
                if (dim[i] > maxd) maxd = dim[i];
                for (int j = 0; j <= i; j++, ij++) { ish[ij] = i; jsh[ij] = j; }
        }
        void *opt = NULL;
        optim(&opt, atm, natm, bas, nbas, env);
        double nints = 0, checksum = 0;
        long nquart = 0;
        double t0 = omp_get_wtime();
#pragma omp parallel reduction(+ : nints, checksum, nquart)
        {
                double *buf = malloc(sizeof(double) * maxd * maxd * maxd * maxd);
#pragma omp for schedule(dynamic, 2)
                for (long p = phase; p < npair; p += stride) {
                        int i = ish[p], j = jsh[p];
                        long klmax = (long)(i + 1) * (i + 2) / 2;
                        for (long kl = 0; kl < klmax; kl++) {
                                int shls[4] = {i, j, ish[kl], jsh[kl]};
                                intor(buf, NULL, shls, atm, natm, bas, nbas, env, opt, NULL);
                                long n = (long)dim[i] * dim[j] * dim[shls[2]] * dim[shls[3]];
                                nints += n;
                                checksum += buf[0] + buf[n - 1];
                        }
                        nquart += klmax;
                }
                free(buf);
        }
        double t1 = omp_get_wtime();
        if (delopt) delopt(&opt);
        printf("{\"seconds\": %.6f, \"integrals\": %.0f, \"quartets\": %ld, \"threads\": %d, \"stride\": %ld, \"phase\": %ld, \"checksum\": %.15e}\n",
               t1 - t0, nints, nquart, omp_get_max_threads(), stride, phase, checksum);
        return 0;
}
