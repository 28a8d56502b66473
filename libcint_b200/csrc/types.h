// Shared host/device data layout of the B200 ERI engine.
//
// HBM layout (all built once per context, read-only afterwards):
//   PairHdr  pairs[npair]     one per unordered shell pair {i,j} (canonical orientation: the shell
//                             with the larger angular momentum is `a`), followed by nbas
//                             single-shell pseudo pairs used as the ket of 3-centre integrals.
//   PrimPair prims[]          surviving primitive pairs of every shell pair, back to back (64 B each).
//   double   pcoef[]          contraction-coefficient products c_a[ca][ip] * c_b[cb][jp] per
//                             primitive pair, nca*ncb doubles each (ca fastest).
// This is the device counterpart of the reference's CINTOpt/PairData (include/cint.h.in:140-152,
// src/optimizer.c:288-342); screening uses the same cce estimate.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include <stddef.h>

#define B200_LMAX 6            // per-shell angular momentum limit of this build
#define B200_MAXROOTS 13       // (4*LMAX)/2+1
#define B200_NCART(l) (((l) + 1) * ((l) + 2) / 2)

struct PairHdr {               // 96 bytes
    int sh_a, sh_b;            // shell ids in canonical orientation (sh_b = -1: single-shell pseudo pair)
    int la, lb;
    int npp;                   // surviving primitive pairs
    int pp_off;                // first PrimPair
    int nca, ncb;              // contraction counts of a, b
    int cc_off;                // first coefficient product (doubles)
    int ao_a, ao_b;            // AO offsets (spherical) of a and b -- used by tile mode only
    int pad;
    double ra[3];              // centre of a
    double ab[3];              // ra - rb  (HRR shift)
};

struct PrimPair {              // 64 bytes, 16-byte aligned for vector loads
    double aij;                // a_a + a_b
    double px, py, pz;         // Gaussian product centre
    double kij;                // exp(-a_a a_b / aij |ab|^2) (times 2 sqrt(pi) for pseudo pairs)
    double cce;                // screening estimate of CINTset_pairdata (src/optimizer.c:325)
    double inv_aij;            // 1 / aij
    int ipa, ipb;              // primitive indices inside shells a, b
};

// One unit of work for the kernels: (bra pair | ket pair) -> block written at out + off with
// strides given for the CANONICAL indices a,b,c,d (elements).
struct Task {
    int bra, ket;              // indices into pairs[]
    int sb;                    // stride of b   (stride of a is sa)
    int sa;
    long long sc, sd;          // strides of c, d
    long long off;             // block offset in `out`
    int flags;                 // bits 0..3: canonical index a/b/c/d stays Cartesian although the call is spherical
    int pad;                   //            (first-derivative assembly needs the differentiated index in Cartesians)
};

struct EngineParams {          // passed by value to kernels
    const PairHdr *pairs;
    const PrimPair *prims;
    const double *pcoef;
    const double *rys_coef;    // all nroots tables back to back (RYS_TAB_COEF)
    const double *c2s;         // C2S_COEF
    double expcutoff;          // quartet primitive screening: cce_ij + cce_kl <= expcutoff
    double omega;              // env[PTR_RANGE_OMEGA]; > 0 long range, 0 plain Coulomb
    int cart;                  // 1: Cartesian output (int2e_cart), 0: real spherical
};

// ----------------------------------------------------------------------------- tile description
struct TileParams {
    // T side (per-thread pairs): structure-of-arrays tables of one pair class
    const double *tprim;       // [6 + NCT][Q][NT]: aij, 1/aij, px, py, pz, kij, cc[NCT]
    const double *tgeom;       // [6][NT]: ra[3], ab[3]
    const long long *trow;     // [NT] global row offset of the pair's block
    const int *tstride;        // [2][NT] sa, sb (elements)
    const int *tI;             // [NT] larger shell index of the pair
    const int *tpair;          // [NT] pair ids (for the generic kernel's tile mode)
    const double *tq;          // [NT] Schwarz bound sqrt(max|(ab|ab)|) of every T pair (NULL: screening off)
    const double *uq;          // [NU_all] same for the kets
    double schwarz_thr;        // skip a work item when bound(T) * bound(U) < thr for all of its quartets
    const int *tnpp;           // [NT] primitive pairs per T pair (>= 1); lists are sorted by descending count inside a chunk
    int NT, Q;                 // pairs in class, primitives per pair (padded)
    int t_begin, t_end;        // range of this chunk inside the class list
    // list mode (engine.cu:run_batch fast path): explicit work items {ket index u, first T occurrence, valid count, -} and an
    // indirection from T occurrences to the rows of the class' pair table; NULL / 0 in tile mode
    const int4 *items;
    long long nitems;
    const int *tsel;           // [occurrences] row of tprim / tgeom / tnpp used by occurrence t (trow / tstride are per occurrence)
    int NTs;                   // length of one row of tstride (= NT in tile mode)
    // range-separated Coulomb (kernels instantiated with RS = true; env[PTR_RANGE_OMEGA] != 0, src/g2e.c:4443-4492):
    // pass 1 = erf-attenuated rule at theta x with t^2 -> theta t^2 and weight rs_sign sqrt(theta); pass 0 = full Coulomb.
    // long range (omega > 0): pass 1 only, sign +1;  short range (omega < 0): both passes, sign -1 (erfc = 1 - erf)
    double rs_w2, rs_sign;     // omega^2 and sign * |omega| (theta = w2 r^2, sign sqrt(theta) = rs_sign r with r = rsqrt(w2 + a0))
    int rs_pass0;
    unsigned int *counter;     // per-launch work-item counter (zeroed before every job)
    int batch;                 // work items fetched per atomic (sized on the host so that every launch has >= ~8 batches per SM)
    int gx;                    // work items per ket = ceil((t_end - t_begin) / pairs per block)
    int nca_t;                 // contraction count of shell a (T side), for block offsets
    // U side (block-uniform pairs)
    const int *upair;          // [NU] pair ids (AoS tables in EngineParams), sorted by larger shell index
    const int *uK;             // [NU] larger shell index
    const long long *ucol;     // [NU] column offset of the pair's block (this rank's numbering)
    const int *ustride;        // [2][NU] sc, sd (in columns)
    int NU;                    // kets of this launch (this rank's share among the first NU_valid)
    int NU_all;                // length of the class list (second row of ustride starts here)
    int u_step, u_first;       // rank sharding: u = u_first + u_step * blockIdx.y
    int nca_u;
    int umax;                  // most primitive pairs of any ket of this launch (sizes the per-warp smem slice)
    int tri;                   // 0: every T pair of [t_begin, t_end) meets every ket; 1: only quartets with K(u) <= I(t) (reference
                               // benchmark loop; the range lies in the shell-sorted ordering B); 2: kets with K < tri_i0 take the range
                               // as it is (ordering A), the others the suffix of [t_begin + tB, t_end + tB) with I >= K (ordering B)
    int tB, tri_i0;
    // output tile
    double *out;
    long long row0;            // global row of the chunk's first row
    long long ld;              // rows of the chunk buffer
    // engine tables
    const PairHdr *pairs;
    const PrimPair *prims;
    const double *pcoef;
    const double *rys;         // table of this class' nroots
};
