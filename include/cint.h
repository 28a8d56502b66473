/*
 * libcint-compatible C ABI for the ERI hot path, served by the B200 (sm_100a) implementation.
 *
 * This header is written for this repository; it declares, with the reference's names, argument
 * order and meaning, exactly the entry points of the reference that belong to the hot path
 * (SURVEY.md section 8b).  Each declaration cites the reference interface it replaces.
 * Everything executes on the GPU: there is no CPU fallback, and every call fails loudly
 * (message on stderr, return value 0 / NULL optimizer) if no CUDA device is usable.
 *
 * Array layouts (reference: include/cint.h.in:28-67, doc/program_ref.txt:30-70):
 *   atm[natm][ATM_SLOTS] int, bas[nbas][BAS_SLOTS] int, env[] double; env[0..19] reserved:
 *   env[PTR_EXPCUTOFF] screening exponent (0 -> default 60), env[PTR_RANGE_OMEGA] range separation.
 *   Contraction coefficients are expected pre-multiplied by CINTgto_norm(l, exponent).
 */
#ifndef CINT_B200_COMPAT_H
#define CINT_B200_COMPAT_H

#define CINT_VERSION   "6.1.1-b200"
#define FINT int
#define CACHE_SIZE_T FINT

/* env slots (include/cint.h.in:28-45) */
#define PTR_EXPCUTOFF           0
#define PTR_COMMON_ORIG         1
#define PTR_RINV_ORIG           4
#define PTR_RINV_ZETA           7
#define PTR_RANGE_OMEGA         8
#define PTR_F12_ZETA            9
#define PTR_GTG_ZETA            10
#define NGRIDS                  11
#define PTR_GRIDS               12
#define PTR_ENV_START           20
/* atm slots (include/cint.h.in:49-56) */
#define CHARGE_OF       0
#define PTR_COORD       1
#define NUC_MOD_OF      2
#define PTR_ZETA        3
#define PTR_FRAC_CHARGE 4
#define RESERVE_ATMSLOT 5
#define ATM_SLOTS       6
/* bas slots (include/cint.h.in:59-67) */
#define ATOM_OF         0
#define ANG_OF          1
#define NPRIM_OF        2
#define NCTR_OF         3
#define KAPPA_OF        4
#define PTR_EXP         5
#define PTR_COEFF       6
#define RESERVE_BASLOT  7
#define BAS_SLOTS       8

#define bas(SLOT,I)     bas[BAS_SLOTS * (I) + (SLOT)]
#define atm(SLOT,I)     atm[ATM_SLOTS * (I) + (SLOT)]

#ifdef __cplusplus
extern "C" {
#endif

/* Opaque here: owns the device-resident basis, shell-pair tables and streams built from one
 * (atm, bas, env).  Reference: public struct CINTOpt, include/cint.h.in:140-152; callers of the hot
 * path only pass it back (examples/time_c60.c:193-217). */
typedef struct CINTOpt CINTOpt;

/* include/cint_funcs.h:11-16 */
typedef CACHE_SIZE_T CINTIntegralFunction(double *out, FINT *dims, FINT *shls,
                                          FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env,
                                          CINTOpt *opt, double *cache);
typedef void CINTOptimizerFunction(CINTOpt **opt, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env);

/* ---- 4-centre ERIs: src/cint2e.c:1186-1210 ---- */
CINTIntegralFunction  int2e_sph;            /* src/cint2e.c:1186 */
CINTIntegralFunction  int2e_cart;           /* src/cint2e.c:1202 */
CINTOptimizerFunction int2e_optimizer;      /* src/cint2e.c:1195 */
/* ---- 3-centre ERIs: src/cint3c2e.c:693-717 (exported by the reference, absent from cint_funcs.h) ---- */
CINTIntegralFunction  int3c2e_sph;          /* src/cint3c2e.c:693 */
CINTIntegralFunction  int3c2e_cart;         /* src/cint3c2e.c:710 */
CINTOptimizerFunction int3c2e_optimizer;    /* src/cint3c2e.c:702 */
CINTIntegralFunction  int3c2e_sph_ssc;      /* src/cint3c2e.c:729: spherical i, j, Cartesian auxiliary index (c2s_sph_3c2e1_ssc, src/cart2sph.c:5956) */
CINTOptimizerFunction int3c2e_ssc_optimizer; /* src/cint3c2e.c:759 */

/* ---- 2-centre ERIs (density-fitting metric): src/cint2c2e.c:351-375 ---- */
CINTIntegralFunction  int2c2e_sph;          /* src/cint2c2e.c:351 */
CINTIntegralFunction  int2c2e_cart;         /* src/cint2c2e.c:368 */
CINTOptimizerFunction int2c2e_optimizer;    /* src/cint2c2e.c:360 */

/* ---- first derivatives ( nabla i j | k l ), 3 components, out[comp][l][k][j][i] ---- */
CINTIntegralFunction  int2e_ip1_sph;        /* src/autocode/grad2.c:51 */
CINTIntegralFunction  int2e_ip1_cart;       /* src/autocode/grad2.c:42 */
CINTOptimizerFunction int2e_ip1_optimizer;  /* src/autocode/grad2.c:35 */
CINTIntegralFunction  int3c2e_ip1_sph;      /* src/autocode/int3c2e.c (int3c2e_ip1_sph) */
CINTIntegralFunction  int3c2e_ip1_cart;
CINTOptimizerFunction int3c2e_ip1_optimizer;
CINTIntegralFunction  int3c2e_ip2_sph;      /* ( i j | nabla k ), src/autocode/int3c2e.c:161 */
CINTIntegralFunction  int3c2e_ip2_cart;     /* src/autocode/int3c2e.c:153 */
CINTOptimizerFunction int3c2e_ip2_optimizer;
CINTIntegralFunction  int2c2e_ip1_sph;      /* ( nabla i | k ), src/autocode/int3c2e.c:376 */
CINTOptimizerFunction int2c2e_ip1_optimizer;
CINTIntegralFunction  int2c2e_ip2_sph;      /* ( i | nabla k ), src/autocode/int3c2e.c:454 */
CINTOptimizerFunction int2c2e_ip2_optimizer;

/* ---- v2-style wrappers (src/misc.h:35-61 ALL_CINT, include/cint.h.in:264-278) ---- */
FINT cint2e_sph(double *out, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt);
FINT cint2e_cart(double *out, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt);
void cint2e_sph_optimizer(CINTOpt **opt, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env);
void cint2e_cart_optimizer(CINTOpt **opt, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env);
FINT cint3c2e_sph(double *out, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt);
FINT cint3c2e_cart(double *out, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt);
void cint3c2e_sph_optimizer(CINTOpt **opt, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env);
void cint3c2e_cart_optimizer(CINTOpt **opt, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env);
FINT cint2e_ip1_sph(double *out, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt);
FINT cint3c2e_ip1_sph(double *out, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt);
FINT cint3c2e_ip2_sph(double *out, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt);
FINT cint2c2e_ip1_sph(double *out, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt);
FINT cint2c2e_ip2_sph(double *out, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt);
void cint2e_ip1_sph_optimizer(CINTOpt **opt, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env);
void cint3c2e_ip1_sph_optimizer(CINTOpt **opt, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env);
FINT cint2c2e_sph(double *out, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt);
FINT cint2c2e_cart(double *out, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt);
void cint2c2e_sph_optimizer(CINTOpt **opt, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env);
void cint2c2e_cart_optimizer(CINTOpt **opt, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env);

/* ---- optimizer life cycle: src/optimizer.c:22-72 ---- */
void CINTinit_2e_optimizer(CINTOpt **opt, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env);
void CINTinit_optimizer(CINTOpt **opt, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env);
void CINTdel_2e_optimizer(CINTOpt **opt);
void CINTdel_optimizer(CINTOpt **opt);

/* ---- shell bookkeeping used by the benchmark drivers: src/cint_bas.c ---- */
FINT CINTlen_cart(const FINT l);                                        /* src/cint_bas.c:12 */
FINT CINTcgtos_cart(const FINT bas_id, const FINT *bas);                /* src/cint_bas.c:31 */
FINT CINTcgto_cart(const FINT bas_id, const FINT *bas);                 /* src/cint_bas.c:36 */
FINT CINTcgtos_spheric(const FINT bas_id, const FINT *bas);             /* src/cint_bas.c:45 */
FINT CINTcgto_spheric(const FINT bas_id, const FINT *bas);              /* src/cint_bas.c:49 */
FINT CINTtot_pgto_spheric(const FINT *bas, const FINT nbas);            /* src/cint_bas.c:69 */
FINT CINTtot_cgto_spheric(const FINT *bas, const FINT nbas);            /* src/cint_bas.c:108 */
FINT CINTtot_cgto_cart(const FINT *bas, const FINT nbas);               /* src/cint_bas.c:124 */
void CINTshells_cart_offset(FINT ao_loc[], const FINT *bas, const FINT nbas);       /* src/cint_bas.c:141 */
void CINTshells_spheric_offset(FINT ao_loc[], const FINT *bas, const FINT nbas);    /* src/cint_bas.c:149 */
double CINTgto_norm(FINT n, double a);                                  /* src/misc.c:86 */

#ifdef __cplusplus
}
#endif
#endif
