/*
 * Fixture dumper (test tooling, not product code).
 *
 * The reference's benchmark drivers (examples/time_c60.c, examples/time_c2h6.c) build their
 * atm/bas/env arrays inline and hand them to the library.  To obtain exactly those arrays without
 * copying the drivers, the UNMODIFIED example is compiled against this shim instead of libcint:
 * the shim implements the handful of symbols the drivers call, records the (atm, bas, env) passed
 * to cint2e_sph_optimizer() into $FIXTURE_OUT.<n>.bin, and makes the integral calls no-ops so the
 * driver's loops finish immediately.  tools/make_fixtures.py turns the dumps into .npz files.
 *
 * CINTgto_norm restates the closed form of src/misc.c:86 (radial normalisation of r^l exp(-a r^2)):
 *   N = 1/sqrt( int_0^inf r^(2l+2) exp(-2 a r^2) dr ),  with  int = Gamma(l+1.5) / (2 (2a)^(l+1.5)).
 */
#include <stdio.h>
#include <stdlib.h>
#include <math.h>

#define ATM_SLOTS 6
#define BAS_SLOTS 8
typedef struct CINTOpt CINTOpt;

double CINTgto_norm(int n, double a)
{
        double p = n + 1.5;
        return 1.0 / sqrt(tgamma(p) / (2.0 * pow(2.0 * a, p)));
}

static int ncart_or_sph(const int *bas, int ib) { return (2 * bas[ib*BAS_SLOTS+1] + 1) * bas[ib*BAS_SLOTS+3]; }
int CINTcgto_spheric(int ib, const int *bas) { return ncart_or_sph(bas, ib); }
int CINTtot_cgto_spheric(const int *bas, int nbas)
{
        int i, s = 0;
        for (i = 0; i < nbas; i++) s += ncart_or_sph(bas, i);
        return s;
}
int CINTtot_pgto_spheric(const int *bas, int nbas)
{
        int i, s = 0;
        for (i = 0; i < nbas; i++) s += (2 * bas[i*BAS_SLOTS+1] + 1) * bas[i*BAS_SLOTS+2];
        return s;
}

static int ndump = 0;
void cint2e_sph_optimizer(CINTOpt **opt, int *atm, int natm, int *bas, int nbas, double *env)
{
        const char *base = getenv("FIXTURE_OUT");
        char name[512];
        int i, nenv = 20;
        for (i = 0; i < natm; i++)
                if (atm[i*ATM_SLOTS+1] + 3 > nenv) nenv = atm[i*ATM_SLOTS+1] + 3;
        for (i = 0; i < nbas; i++) {
                int np = bas[i*BAS_SLOTS+2], nc = bas[i*BAS_SLOTS+3];
                if (bas[i*BAS_SLOTS+5] + np > nenv) nenv = bas[i*BAS_SLOTS+5] + np;
                if (bas[i*BAS_SLOTS+6] + np*nc > nenv) nenv = bas[i*BAS_SLOTS+6] + np*nc;
        }
        snprintf(name, sizeof name, "%s.%d.bin", base ? base : "fixture", ndump++);
        FILE *f = fopen(name, "wb");
        int hdr[3] = {natm, nbas, nenv};
        fwrite(hdr, sizeof(int), 3, f);
        fwrite(atm, sizeof(int), (size_t)natm*ATM_SLOTS, f);
        fwrite(bas, sizeof(int), (size_t)nbas*BAS_SLOTS, f);
        fwrite(env, sizeof(double), nenv, f);
        fclose(f);
        *opt = NULL;
}
void cint2e_ip1_sph_optimizer(CINTOpt **opt, int *atm, int natm, int *bas, int nbas, double *env) { *opt = NULL; }
int cint2e_sph(double *buf, int *shls, int *atm, int natm, int *bas, int nbas, double *env, CINTOpt *opt) { return 0; }
int cint2e_ip1_sph(double *buf, int *shls, int *atm, int natm, int *bas, int nbas, double *env, CINTOpt *opt) { return 0; }
void CINTdel_optimizer(CINTOpt **opt) { *opt = NULL; }
