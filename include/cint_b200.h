/*
 * Batched shell-tuple entry points of the B200 ERI engine (new in this repository; the reference
 * has no batched call -- its callers loop over shell quartets from OpenMP threads,
 * examples/time_c60.c:196-219).  Plain C ABI: pointers, ints and size_t only.
 *
 * A context owns everything that depends on one (atm, bas, env): the device copy of the basis,
 * the screened primitive-pair tables (the device counterpart of CINTOpt's PairData,
 * src/optimizer.c:288-342) and class-sorted shell-pair lists.  It is the same object an
 * `*_optimizer` call of include/cint.h returns as CINTOpt.
 *
 * Every function returns a negative CINTB200_E* code on failure and prints the reason to stderr.
 * There is no CPU fallback.
 */
#ifndef CINT_B200_H
#define CINT_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct CINTOpt cintb200_ctx;   /* same object as the CINTOpt of include/cint.h */

#define CINTB200_SPH   0
#define CINTB200_CART  1

#define CINTB200_ENODEV   (-1)   /* no usable CUDA device / CUDA runtime error */
#define CINTB200_EINVAL   (-2)   /* bad argument (shell id, angular momentum beyond CINTB200_LMAX, ...) */
#define CINTB200_ENOMEM   (-3)
#define CINTB200_ENOSUP   (-4)   /* valid libcint input this build does not implement yet */
#define CINTB200_LMAX     6      /* highest angular momentum accepted per shell */

/* Build / destroy a context on CUDA device `device` (-1: current device).  Replaces the work of
 * CINTall_2e_optimizer (src/optimizer.c:183) + the per-call CINTinit_int2e_EnvVars (src/g2e.c:21). */
int  cintb200_create(cintb200_ctx **ctx, const int *atm, int natm, const int *bas, int nbas,
                     const double *env, int device);
void cintb200_destroy(cintb200_ctx *ctx);
int  cintb200_device(const cintb200_ctx *ctx);

/*
 * Evaluate n shell quartets (ij|kl), shls[4*t .. 4*t+3] = i,j,k,l (0-based shell ids).
 *   kind      CINTB200_SPH  -> values of int2e_sph  (src/cint2e.c:1186)
 *             CINTB200_CART -> values of int2e_cart (src/cint2e.c:1202)
 *   out_off   element offset of each quartet's block inside `out`; NULL -> blocks packed back to back
 *             in input order.  Each block is column-major (di,dj,dk,dl), i fastest, exactly the
 *             `out` of the reference call with dims == NULL.
 *   out       host pointer (on_device = 0; staged through pinned memory) or device pointer
 *             (on_device = 1) with room for every block.
 *   nonzero   optional host array[n]: the reference's return value per quartet (1 if any primitive
 *             survived exponent screening, else 0 and the block is zero-filled).
 * Returns the number of quartets evaluated (n) or a negative error code.
 */
long cintb200_int2e_batch(cintb200_ctx *ctx, int kind, const int *shls, size_t n,
                          const size_t *out_off, double *out, int on_device, int *nonzero);

/* Same for shell triples (ij|k), shls[3*t .. 3*t+2]; block (di,dj,dk).  int3c2e_sph/_cart,
 * src/cint3c2e.c:693,710. */
long cintb200_int3c2e_batch(cintb200_ctx *ctx, int kind, const int *shls, size_t n,
                            const size_t *out_off, double *out, int on_device, int *nonzero);

/* Same for shell pairs (i|k), shls[2*t .. 2*t+1]; block (di,dk): the density-fitting metric.  int2c2e_sph/_cart,
 * src/cint2c2e.c:351,368 (same Rys engine with aj = al = 0, src/g2c2e.c:15-100). */
long cintb200_int2c2e_batch(cintb200_ctx *ctx, int kind, const int *shls, size_t n,
                            const size_t *out_off, double *out, int on_device, int *nonzero);

/* First derivatives ( nabla i j | k l ) and ( nabla i j | k ): int2e_ip1_sph/_cart (src/autocode/grad2.c:19-68) and
 * int3c2e_ip1_sph/_cart.  Every block holds 3 components, out[comp][l][k][j][i] -- 3x the size of the plain block, the
 * reference's layout for dims == NULL.  Evaluated as a combination of the blocks of a raised and a lowered shell i
 * (the shell-level form of CINTnabla1i_2e, src/g2e.c:4550) through the generic kernel + one assembly kernel. */
long cintb200_int2e_ip1_batch(cintb200_ctx *ctx, int kind, const int *shls, size_t n,
                              const size_t *out_off, double *out, int on_device, int *nonzero);
long cintb200_int3c2e_ip1_batch(cintb200_ctx *ctx, int kind, const int *shls, size_t n,
                                const size_t *out_off, double *out, int on_device, int *nonzero);
/* ( i j | nabla k ) -- the auxiliary-centre gradient of density fitting -- and the 2-centre ( nabla i | k ), ( i | nabla k ):
 * int3c2e_ip2, int2c2e_ip1, int2c2e_ip2 (src/autocode/int3c2e.c:99-168, :330-383, :408-461); same layout, 3 components. */
long cintb200_int3c2e_ip2_batch(cintb200_ctx *ctx, int kind, const int *shls, size_t n,
                                const size_t *out_off, double *out, int on_device, int *nonzero);
long cintb200_int2c2e_ip1_batch(cintb200_ctx *ctx, int kind, const int *shls, size_t n,
                                const size_t *out_off, double *out, int on_device, int *nonzero);
long cintb200_int2c2e_ip2_batch(cintb200_ctx *ctx, int kind, const int *shls, size_t n,
                                const size_t *out_off, double *out, int on_device, int *nonzero);

/* Size in doubles of one block / of a packed batch (host-side helper). */
size_t cintb200_block_size(const cintb200_ctx *ctx, int kind, const int *shls, int ncenter);

/*
 * Whole-job driver for the reference benchmark loop (examples/time_c60.c:200-219): every unique
 * quartet i>=j, k>=l, k<=i of int2e_sph, evaluated class by class into a device-resident ring of
 * `chunk_bytes` (0 -> default).  Rank `rank` of `nranks` evaluates a static, cost-balanced shard
 * (no communication).  If `host_sink` != NULL every finished chunk is copied to it (pinned host
 * buffer of at least chunk_bytes) inside the call -- the end-to-end mode.
 *   stats[0] shell quartets evaluated     stats[1] integrals written
 *   stats[2] primitive quartets executed  stats[3] sum of all integrals (with checksums on, see below)
 *   stats[4] kernel launches              stats[5] device->host bytes
 *   stats[6] model FLOPs of the executed primitive quartets (SURVEY 8d formula)
 *   stats[7] GPU milliseconds of the ERI kernels (CUDA events on the launch stream)
 */
int cintb200_int2e_sph_all_unique(cintb200_ctx *ctx, int rank, int nranks, size_t chunk_bytes,
                                  double *host_sink, double *stats);
/*
 * Notes on the whole-job drivers:
 *  - stats[3] is filled when checksums are switched on (cintb200_set_checksums), else it is 0.
 *  - chunk boundaries do not depend on nranks: every rank count walks the same chunks with 1/nranks of the columns, so a
 *    rank's tiles are about chunk_bytes / nranks.
 *  - a single `host_sink` is a throughput mode: every tile is copied to the SAME buffer, so after the call only the last
 *    tile can be read.  To CONSUME every tile use the *_tiles variants below (ring of sinks + callback).
 *  - tile entries of quartets outside the loop (K > I, only possible for the kets of the chunk's own bra shells, columns
 *    >= cintb200_tile.ncols_below) are zero in host sinks; in the device-resident tile they are left unwritten.
 */

/* One finished tile, column-major values[row + nrows * col]: rows = AO pairs of the bra shell pairs with i in [i0, i1)
 * (global row numbers row0 .. row0 + nrows), columns = AO pairs of this rank's kets with k < i1 (the first ncols columns of
 * the rank's column numbering; the first ncols_below of them belong to kets with k < i0, for which every row is in the
 * loop).  cintb200_job_row_map / cintb200_job_col_map translate rows and columns to shells and block positions. */
typedef struct {
    int chunk, nchunks, rank, nranks, i0, i1;
    long long row0, nrows, ncols, ncols_below;
} cintb200_tile;
/* Called on the thread that made the whole-job call, once per tile, in chunk order, after the tile has arrived in host memory
 * and before its sink is reused; `values` points into one of the caller's sinks.  Return non-zero to abort the job. */
typedef int (*cintb200_tile_fn)(void *user, const cintb200_tile *tile, const double *values);

/* cintb200_int2e_sph_all_unique / cintb200_int3c2e_sph_all with every tile delivered to the caller: tile k is copied into
 * sinks[k % nsinks] (pinned host buffers of at least chunk_bytes each) while the kernels of tile k+1 run, then `fn` is called
 * (fn may be NULL: copies only).  nsinks >= 2 lets copy, kernels and the consumer overlap; stats[5] = bytes copied. */
int cintb200_int2e_sph_all_unique_tiles(cintb200_ctx *ctx, int rank, int nranks, size_t chunk_bytes, double *const *sinks, int nsinks,
                                        cintb200_tile_fn fn, void *user, double *stats);
int cintb200_int3c2e_sph_all_tiles(cintb200_ctx *ctx, int aux_shell0, int rank, int nranks, size_t chunk_bytes, double *const *sinks,
                                   int nsinks, cintb200_tile_fn fn, void *user, double *stats);

/*
 * Coulomb and exchange matrices built on the device from the tiles of the whole int2e_sph job -- the consumer the reference's
 * callers wrap around the per-quartet call; no integral leaves the GPU:
 *     vj[a,b] = sum_cd (ab|cd) dm[c,d]        vk[a,c] = sum_bd (ab|cd) dm[b,d]
 * dm: SYMMETRIC nao x nao density (spherical AOs in shell order); vj, vk: nao x nao (either may be NULL).  Host pointers
 * (on_device = 0) or device pointers (on_device = 1).  Only the 8-fold unique quartets (ij >= kl) are digested, with the
 * usual symmetry factors; rank `rank` of `nranks` digests its ket shard and returns PARTIAL matrices -- the caller adds them
 * over the ranks (one all-reduce of 2 nao^2 doubles; the only collective of the path).  stats as for the whole-job driver.
 */
int cintb200_int2e_sph_jk(cintb200_ctx *ctx, int rank, int nranks, size_t chunk_bytes, const double *dm, double *vj, double *vk,
                          int on_device, double *stats);

/*
 * Whole-job checksums.  When switched on, every finished tile is reduced on the device to per-row sums of its in-loop entries,
 * and stats[3] returns the sum of all integrals.  cintb200_job_checksums returns, per bra shell pair p = i(i+1)/2 + j, this
 * rank's partial sums over its kets of
 *     S[p] = sum v,   A[p] = sum |v|,   F[p] = sum v cos(0.91 r + 0.3) cos(0.37 c + 0.61 d + 0.5)
 * (r = position mi + di mj of the row inside the (i,j) block, c / d = global spherical AO indices of the ket functions;
 * 3-centre jobs: c = index of the auxiliary function counted from the first auxiliary AO, d = 0).  Summed over the ranks
 * they are independent of chunking and sharding -- the quantities oracle/ref_golden.c computes from the reference.
 * Any of S, A, F may be NULL; returns the number of pairs.
 */
int cintb200_set_checksums(cintb200_ctx *ctx, int on);
int cintb200_job_checksums(cintb200_ctx *ctx, double *S, double *A, double *F);

/* Geometry of the cached whole-job plan (after any whole-job call):
 * geom[0..8] = i0, i1, row0, nrows, ncols, number of chunks, total rows, this rank's total columns, ncols_below. */
int cintb200_job_geometry(cintb200_ctx *ctx, int chunk, long long *geom);
/* per row of chunk `chunk`: bra shells i >= j and the position mi + di mj inside the (i,j) block (arrays of nrows ints, may be NULL) */
int cintb200_job_row_map(cintb200_ctx *ctx, int chunk, int *sh_i, int *sh_j, int *pos);
/* per column: ket shells k >= l (3-centre jobs: auxiliary shell k, l = -1) and the position mk + dk ml (arrays of ncols ints) */
int cintb200_job_col_map(cintb200_ctx *ctx, int chunk, int *sh_k, int *sh_l, int *pos);

/*
 * Whole-job driver for density fitting: every shell triple (ij|k) of int3c2e_sph (src/cint3c2e.c:693) with orbital
 * shells i >= j in [0, aux_shell0) and auxiliary shells k in [aux_shell0, nbas) -- orbital and auxiliary basis share one
 * bas array, as in the reference's callers.  Output tiles are column-major out[row(ij) + ld * col(k)]: a row block is
 * the (di, dj) block of the shell pair laid out like the reference's buf (i fastest), a column block the dk functions
 * of shell k.  Rows are streamed through the device buffer chunk by chunk (ranges of i); the auxiliary shells are dealt
 * round-robin to the ranks inside every (l, nctr) class (static sharding, no communication).  host_sink / stats as above.
 */
int cintb200_int3c2e_sph_all(cintb200_ctx *ctx, int aux_shell0, int rank, int nranks, size_t chunk_bytes,
                             double *host_sink, double *stats);

/*
 * Dense sub-tensor over shell slices -- the shape in which the reference's callers consume integrals (pyscf's fill drivers
 * loop over shell slices and call the per-quartet function for every combination):
 *   int2e:   shls_slice = {i0,i1, j0,j1, k0,k1, l0,l1},  out[i + NI (j + NJ (k + NK l))]  (NI*NJ*NK*NL doubles)
 *   int3c2e: shls_slice = {i0,i1, j0,j1, k0,k1},         out[i + NI (j + NJ k)]
 * AO indices are relative to the slice starts, column-major, spherical; EVERY combination of the slices is evaluated (no
 * permutational symmetry is assumed -- pass triangular slices to exploit it).  Runs on the specialised tile kernels
 * (generic kernel only for classes without one).  out: host pointer (on_device = 0) or device pointer (on_device = 1).
 * stats as for cintb200_int2e_sph_all_unique (may be NULL).
 */
int cintb200_int2e_sph_block(cintb200_ctx *ctx, const int *shls_slice, double *out, int on_device, double *stats);
int cintb200_int3c2e_sph_block(cintb200_ctx *ctx, const int *shls_slice, double *out, int on_device, double *stats);
/* 2-centre metric (i|k): shls_slice = {i0,i1, k0,k1}, out[i + NI k] (src/cint2c2e.c:351). */
int cintb200_int2c2e_sph_block(cintb200_ctx *ctx, const int *shls_slice, double *out, int on_device, double *stats);

/* Cartesian output (int2e_cart src/cint2e.c:1202, int3c2e_cart src/cint3c2e.c:710, int2c2e_cart src/cint2c2e.c:368) from the same
 * specialised kernels, the cart->sph stages compiled out: block dimensions and AO offsets are the Cartesian ones
 * ((l+1)(l+2)/2 functions per shell and contraction), everything else as for the spherical calls above. */
int cintb200_int2e_cart_block(cintb200_ctx *ctx, const int *shls_slice, double *out, int on_device, double *stats);
int cintb200_int3c2e_cart_block(cintb200_ctx *ctx, const int *shls_slice, double *out, int on_device, double *stats);
int cintb200_int2c2e_cart_block(cintb200_ctx *ctx, const int *shls_slice, double *out, int on_device, double *stats);
int cintb200_int2e_cart_all_unique(cintb200_ctx *ctx, int rank, int nranks, size_t chunk_bytes, double *host_sink, double *stats);
int cintb200_int2e_cart_all_unique_tiles(cintb200_ctx *ctx, int rank, int nranks, size_t chunk_bytes, double *const *sinks, int nsinks,
                                         cintb200_tile_fn fn, void *user, double *stats);

/* First derivatives on dense shell-slice blocks, evaluated by the specialised kernels (raised / lowered shell blocks in Cartesians,
 * derivative and cart->sph on the dense tensor):  ( nabla i j | k l ), int2e_ip1_sph / _cart (src/autocode/grad2.c:19-68) and
 * ( nabla i j | k ), int3c2e_ip1 (src/autocode/int3c2e.c).  out[i + NI (j + NJ (k + NK (l + NL comp)))], comp = x, y, z: 3x the plain
 * block; kind = CINTB200_SPH or CINTB200_CART; shls_slice as for the plain block calls. */
int cintb200_int2e_ip1_block(cintb200_ctx *ctx, int kind, const int *shls_slice, double *out, int on_device, double *stats);
int cintb200_int3c2e_ip1_block(cintb200_ctx *ctx, int kind, const int *shls_slice, double *out, int on_device, double *stats);

/* Schwarz screening of the whole-job driver: work items (32 quartets) whose bounds sqrt(max|(ij|ij)|) * sqrt(max|(kl|kl)|)
 * are all below `thr` are not evaluated and their blocks are zero-filled.  Default 1e-15 (errors below the 1e-12 parity
 * tolerance by construction); 0 switches it off.  The bounds are evaluated on the device on first use.
 * cintb200_schwarz_bounds copies them (one per shell pair i >= j, index i(i+1)/2 + j; q may be NULL) and returns their number. */
int cintb200_set_schwarz_threshold(cintb200_ctx *ctx, double thr);
int cintb200_schwarz_bounds(cintb200_ctx *ctx, double *q);

/* Host-only planning of the whole job for `rank` of `nranks` (no GPU needed): the static sharding that
 * cintb200_int2e_sph_all_unique will execute.  out[0] shell quartets, out[1] integrals, out[2] primitive quartets,
 * out[3] model FLOPs, out[4] tile columns owned by this rank, out[5] tile rows (all pairs), out[6] chunks,
 * out[7] kernel launches, out[8] bytes of one tile buffer. */
int cintb200_plan_summary(const int *atm, int natm, const int *bas, int nbas, const double *env,
                          int rank, int nranks, size_t chunk_bytes, double *out);

/* Same for the density-fitting job cintb200_int3c2e_sph_all (out[0] counts shell triples, out[4] this rank's auxiliary columns). */
int cintb200_plan_summary_3c(const int *atm, int natm, const int *bas, int nbas, const double *env, int aux_shell0,
                             int rank, int nranks, size_t chunk_bytes, double *out);

/* Measured FP64 FMA peak of the device in TFLOP/s (DFMA-chain microbenchmark run for about `seconds`);
 * the roofline denominator of bench.py, since MEASURED_PEAKS.json has no FP64 entry. */
int cintb200_fp64_peak(int device, double seconds, double *tflops);
/* SMs x 64 FP64 lanes x 2 flops x SM clock (sm_mhz <= 0: the device's maximum clock), for comparison with the measured figure. */
int cintb200_fp64_peak_theoretical(int device, double sm_mhz, double *tflops);

/* Every device buffer of the library (tables, tiles, work arrays) comes from the device's stream-ordered memory pool and stays
 * cached there when a context is destroyed, so that the next context reuses it (freeing and re-mapping tens of GB costs ~1 s).
 * This call returns the cached memory of `device` (-1: current) to the driver, e.g. before another library needs the HBM. */
int cintb200_release_cached_memory(int device);

/* Last error text of the calling thread ("" if none). */
const char *cintb200_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
