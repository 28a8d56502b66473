/*
 * TEST INFRASTRUCTURE: CPU oracle for the ERI hot path (plain-C restatement of the reference
 * algorithm).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 * Parity status: PINNED -- checked against the reference's known-answer values
 * (testsuite/test_cint.py:479, testsuite/test_3c2e.py:303) and element-wise against the reference
 * itself compiled into oracle/_ref (see tests/test_oracle.py).
 */
#ifndef ERI_ORACLE_H
#define ERI_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

int oracle_rys_t2w(int n, double x, double *t2, double *w);
int oracle_rys_roots(int n, double x, double *u, double *w);
int oracle_sr_rys_roots(int n, double x, double lower, double *u, double *w);

/* same argument meaning as the reference's int2e_sph / int2e_cart / int3c2e_sph
 * (include/cint_funcs.h:13-15); opt and cache are accepted and ignored. */
int oracle_int2e_sph(double *out, const int *dims, const int *shls, const int *atm, int natm,
                     const int *bas, int nbas, const double *env);
int oracle_int2e_cart(double *out, const int *dims, const int *shls, const int *atm, int natm,
                      const int *bas, int nbas, const double *env);
int oracle_int3c2e_sph(double *out, const int *dims, const int *shls, const int *atm, int natm,
                       const int *bas, int nbas, const double *env);
int oracle_int3c2e_cart(double *out, const int *dims, const int *shls, const int *atm, int natm,
                        const int *bas, int nbas, const double *env);
/* 2-centre metric (i|k), src/cint2c2e.c:351,368 */
int oracle_int2c2e_sph(double *out, const int *dims, const int *shls, const int *atm, int natm,
                       const int *bas, int nbas, const double *env);
int oracle_int2c2e_cart(double *out, const int *dims, const int *shls, const int *atm, int natm,
                        const int *bas, int nbas, const double *env);
/* first derivative ( nabla i j | k l ), out[comp][l][k][j][i]: src/autocode/grad2.c:19-68, src/autocode/int3c2e.c */
int oracle_int2e_ip1_sph(double *out, const int *dims, const int *shls, const int *atm, int natm,
                         const int *bas, int nbas, const double *env);
int oracle_int2e_ip1_cart(double *out, const int *dims, const int *shls, const int *atm, int natm,
                          const int *bas, int nbas, const double *env);
int oracle_int3c2e_ip1_sph(double *out, const int *dims, const int *shls, const int *atm, int natm,
                           const int *bas, int nbas, const double *env);
int oracle_int3c2e_ip1_cart(double *out, const int *dims, const int *shls, const int *atm, int natm,
                            const int *bas, int nbas, const double *env);
/* ( i j | nabla k ), ( nabla i | k ), ( i | nabla k ): src/autocode/int3c2e.c:99-168, :330-383, :408-461 */
int oracle_int3c2e_ip2_sph(double *out, const int *dims, const int *shls, const int *atm, int natm,
                           const int *bas, int nbas, const double *env);
int oracle_int2c2e_ip1_sph(double *out, const int *dims, const int *shls, const int *atm, int natm,
                           const int *bas, int nbas, const double *env);
int oracle_int2c2e_ip2_sph(double *out, const int *dims, const int *shls, const int *atm, int natm,
                           const int *bas, int nbas, const double *env);
/* real-spherical transformation matrix of one l: c2s[(2l+1)][(l+1)(l+2)/2], row-major */
int oracle_c2s_matrix(int l, double *c2s);
double oracle_gto_norm(int l, double a);

#ifdef __cplusplus
}
#endif
#endif
