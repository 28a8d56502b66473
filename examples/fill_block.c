/*
 * How a C caller of libcint switches a shell-slice fill loop to libcint_b200.
 *
 * Before (the reference's pattern, e.g. examples/time_c60.c:196-219 or pyscf's GTOnr2e_fill_* drivers):
 *     CINTOpt *opt; cint2e_sph_optimizer(&opt, atm, natm, bas, nbas, env);
 *     for every (i, j, k, l) of the slices: cint2e_sph(buf, shls, atm, natm, bas, nbas, env, opt); copy buf into the tensor
 * After: ONE call per slice set; the optimizer object IS the device context.
 *
 * build:  gcc -O2 -Iinclude examples/fill_block.c -Llibcint_b200 -l:libcint_b200.so -lm -o fill_block
 * The arrays below are a toy basis (two s shells and one p shell on two centres); real callers pass their own atm/bas/env.
 */
#include <stdio.h>
#include <stdlib.h>
#include "cint.h"
#include "cint_b200.h"

int main(void)
{
    int atm[2 * 6] = {1, 20, 0, 0, 0, 0, 1, 23, 0, 0, 0, 0};
    int bas[3 * 8] = {0, 0, 1, 1, 0, 26, 27, 0,     /* s on atom 0 */
                      1, 0, 1, 1, 0, 28, 29, 0,     /* s on atom 1 */
                      1, 1, 1, 1, 0, 30, 31, 0};    /* p on atom 1 */
    double env[32] = {0};
    env[20] = 0; env[21] = 0; env[22] = 0;          /* atom 0 */
    env[23] = 0; env[24] = 0; env[25] = 1.4;        /* atom 1 */
    env[26] = 1.2;  env[27] = CINTgto_norm(0, 1.2);
    env[28] = 0.8;  env[29] = CINTgto_norm(0, 0.8);
    env[30] = 0.5;  env[31] = CINTgto_norm(1, 0.5);
    const int natm = 2, nbas = 3;

    CINTOpt *opt = NULL;
    int2e_optimizer(&opt, atm, natm, bas, nbas, env);        /* device context + pair tables */
    if (!opt) { fprintf(stderr, "no CUDA device: %s\n", cintb200_last_error()); return 1; }

    /* the reference's call, unchanged: one quartet */
    int shls[4] = {2, 0, 1, 2};
    double buf[3 * 1 * 1 * 3];
    int nonzero = int2e_sph(buf, NULL, shls, atm, natm, bas, nbas, env, opt, NULL);
    printf("(p s|s p) block: nonzero = %d, first element %.12f\n", nonzero, buf[0]);

    /* the whole 5 x 5 x 5 x 5 tensor in one call */
    int slice[8] = {0, nbas, 0, nbas, 0, nbas, 0, nbas};
    const int nao = CINTtot_cgto_spheric(bas, nbas);
    double *eri = malloc(sizeof(double) * nao * nao * nao * nao);
    if (cintb200_int2e_sph_block(opt, slice, eri, 0 /* host pointer */, NULL) != 0) {
        fprintf(stderr, "block call failed: %s\n", cintb200_last_error());
        return 1;
    }
    printf("(00|00) = %.12f\n", eri[0]);
    free(eri);
    CINTdel_optimizer(&opt);
    return 0;
}
