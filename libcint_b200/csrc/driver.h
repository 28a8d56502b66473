// Whole-job plan of the tile driver (driver.cu) -- shared with the tile consumers (digest.cu: checksums, J/K).
#pragma once
#include <vector>
#include <cuda_runtime.h>
#include "types.h"
#include "kernels.h"
struct CINTOpt;

// Bump allocator for the many small device tables of a plan: a few pooled blocks instead of ~200 cudaMalloc / cudaFree calls
// per plan (each a driver round trip with an implicit synchronisation -- per context, i.e. per end-to-end step).
// Uploads are deferred: every block has a host mirror that upload() fills, and flush() (end of the plan builder) sends each
// block with ONE copy instead of one synchronous cudaMemcpy per table (a dense-block plan has ~150 tables: 2.8 ms per plan before).
struct DeviceArena {
    std::vector<void *> blocks;
    std::vector<char *> mirrors;            // host copies of the blocks (uninitialised storage, only the used part is touched)
    std::vector<size_t> sizes, used, flushed;       // block size; bytes handed out; bytes already sent
    char *cur = nullptr;
    size_t left = 0;
    void *alloc(size_t bytes);
    char *host_of(void *dev);               // mirror address of a device address handed out by alloc()
    int flush();                            // send what has been written since the last flush
    void release();
};

struct PairClass {
    int la, lb, nca, ncb, Q;
    std::vector<int> ids, I, npp;
    std::vector<long long> npp_prefix;      // prefix sums of npp over the list
    double *d_tprim = nullptr, *d_tgeom = nullptr;
    long long *d_trow = nullptr, *d_ucol = nullptr;
    int *d_tstride = nullptr, *d_tI = nullptr, *d_tpair = nullptr, *d_ustride = nullptr, *d_tnpp = nullptr;
    std::vector<int> chunk_lo;              // first list index of every chunk (+ end)
    // second ordering of the same pairs for the DIAGONAL kets of a chunk (K inside the chunk's bra shell range):
    // sorted by the larger shell index (then by primitive count), so the valid bras of a ket are a suffix.  The device
    // tables above hold both orderings back to back (rows [0, NT) and [NT, 2 NT)).
    std::vector<int> idsB, IB, nppB;
    double *d_tq = nullptr;                 // Schwarz bounds, same two orderings
};

struct LaunchRec {
    int chunk;
    TileParams P;               // out / row0 / ld filled per run (buffer alternates)
    RegKernelFn fn;             // nullptr -> generic kernel
    int coop; CoopInfo ci;      // fn is a cooperative kernel
    int nroots, ncu, gx, gy;
    GenericClass GC; GenericLaunch GL;
    long long ntasks;           // generic: number of quartets
    size_t uprefix_off;         // generic: offset into plan->d_uprefix
    int key[6];                 // la lb lc ld nct ncu
    double quartets, prim, flops, integrals;
    int part;                   // 0: kets below the chunk's bra shells, 1: the chunk's own kets
};

struct JobPlan {
    int rank = 0, nranks = 1;
    size_t chunk_bytes = 0;
    int ncenter = 4, aux0 = 0;              // 3: rows = orbital pairs of shells [0, aux0), columns = auxiliary shells [aux0, nbas)
    int rect = 0;                           // dense shell-slice block (build_rect_plan): explicit bra / ket lists, one tile; value =
                                            // number of centres of the integral (3 or 4)
    int cart = 0;                           // Cartesian output (int2e_cart / int3c2e_cart): block dimensions are ncart(l) * nctr
    int own_out = 1;                        // d_out[0] belongs to the plan (rect jobs may write into the caller's device buffer)
    std::vector<PairClass> classes;
    std::vector<PairClass> uclasses;        // 3-centre jobs: classes of the single-shell pseudo pairs (kets); 4-centre: unused
    std::vector<long long> colof_aux;       // 3-centre jobs: this rank's column offset of auxiliary shell aux0 + n, or -1
    std::vector<long long> rowoff;          // per pair id
    std::vector<long long> rows_before;     // [nbas+1] rows of pairs with I < i
    std::vector<long long> cols_before;     // [nbas+1] this rank's columns of kets with K < i
    std::vector<std::pair<int, int>> chunks;
    std::vector<long long> chunk_cols;      // this rank's columns needed by chunk k (kets with K < i1)
    double *d_out[2] = {nullptr, nullptr};
    size_t out_doubles = 0;
    long long *d_uprefix = nullptr; size_t cap_uprefix = 0;
    unsigned int *d_counters = nullptr;     // one work-item counter per launch
    double *d_scratch = nullptr; size_t cap_scratch = 0;
    int force_generic = 0;
    double schwarz_thr = 0;
    int host_only = 0;                      // planning without a device (cintb200_plan_summary)
    std::vector<long long> colof;           // per pair id: this rank's column offset or -1
    std::vector<struct LaunchRec> launches;
    double st_quartets = 0, st_integrals = 0, st_prim = 0, st_flops = 0;
    cudaStream_t copy_stream = nullptr;
#ifndef B200_NSTREAMS
#define B200_NSTREAMS 8
#endif
    static const int NS = B200_NSTREAMS;    // concurrent launch streams (independent classes overlap)
    cudaStream_t streams[NS] = {nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[NS] = {nullptr};
    cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr}, ev_t0 = nullptr, ev_t1 = nullptr;
    // ---- job geometry for consumers of the tiles (digest.cu: checksums, J/K; host callbacks) ----
    std::vector<int> row_pair, row_pos;     // [all rows] bra pair id i(i+1)/2+j of the row and its position mi + di*mj inside the block
    std::vector<int> col_pair, col_pos;     // [this rank's columns] ket pair id (3-centre: auxiliary shell id) and position mk + dk*ml
    DeviceArena arena;                      // owns every table uploaded while the plan was built
    struct DigestState *digest = nullptr;   // device-side state of the tile consumers, built on first use (digest.cu)
    // CUDA graphs of the device-resident job (driver.cu:execute_plan), one per consumer set: [0] plain, [1] with checksums
    struct GraphSlot { cudaGraphExec_t exec = nullptr; int warm = 0; long long nlaunch = 0, reg_launches = 0; };
    GraphSlot graph[2];
};

// ---- list mode on the tile kernels (driver.cu:list_mode_run on the host, listdev.cu on the device) ----
struct ListClass {
    int la, lb, nca, ncb, Q;
    std::vector<int> ids;
    double *d_tprim = nullptr, *d_tgeom = nullptr;
    int *d_tnpp = nullptr;
};
struct ListChoice { RegKernelFn fn; int coop; CoopInfo ci; };
struct ListTables {
    std::vector<ListClass> cls;
    std::vector<int> cls_of, row_of;        // per pair id (shell pairs, then the single-shell pseudo pairs)
    std::vector<ListChoice> choice, choice_cart;   // [bra class * ncls + ket class]: the specialised kernel, if any (spherical / Cartesian output)
    void *d_buf = nullptr; size_t cap = 0;  // per-call arrays (grow-only)
    // device-side list handling (listdev.cu): per pair id class / table row, per shell dimensions {sph, cart}, per group
    // bras per work item (0: no specialised kernel) for spherical and Cartesian output, grow-only work area
    int *d_cls_of = nullptr, *d_row_of = nullptr, *d_sdim = nullptr, *d_per = nullptr;
    void *d_work = nullptr; size_t cap_work = 0;
};

int listtables_build(CINTOpt *c);
int listclass_upload(CINTOpt *c, ListClass &lc);

// ---- tile consumers (digest.cu) ----
struct DigestJob {                          // what to do with every finished tile, besides (optionally) copying it to the host
    int checksums = 0;                      // per-row sums of the valid entries (cintb200_set_checksums)
    int jk = 0;                             // digest into Coulomb / exchange matrices (cintb200_int2e_sph_jk)
    int want_k = 1;
};
void digest_free(struct DigestState *d);
int digest_begin(CINTOpt *c, JobPlan *plan, const DigestJob &job, const double *dm_dev, cudaStream_t st);
int digest_tile(CINTOpt *c, JobPlan *plan, const DigestJob &job, int chunk, const double *tile, cudaStream_t st);
int digest_end(CINTOpt *c, JobPlan *plan, const DigestJob &job, double *vj_dev, double *vk_dev, cudaStream_t st);
void digest_mark_rowsums(JobPlan *plan);     // a replayed graph refreshed the row sums
int digest_fetch_checksums(CINTOpt *c, JobPlan *plan, double *S, double *A, double *F, double *total);


