// Definition of the context object (the CINTOpt handed to callers of include/cint.h).
#pragma once
#include <vector>
#include <mutex>
#include <atomic>
#include <cuda_runtime.h>
#include "types.h"

#define B200_CTX_MAGIC 0x42323030

struct ShellInfo {
    int l, nprim, nctr;
    int ao_sph, ao_cart;
    const double *r, *exps, *coef;      // into CINTOpt::env
};

struct CINTOpt {
    int magic = 0;
    int device = 0;
    cudaStream_t stream = nullptr;
    int natm = 0, nbas = 0;
    int nao_sph = 0, nao_cart = 0;
    uint64_t hash = 0;
    double expcutoff4 = 60, expcutoff3 = 60, omega = 0;
    std::vector<int> atm, bas;
    std::vector<double> env;
    std::vector<ShellInfo> shells;
    std::vector<PairHdr> pairs;
    std::vector<int> vfirst;        // per pair i >= j: first of its nca x ncb VIRTUAL segmented pairs in `pairs` (index cb * nca + ca), or -1
    std::vector<PrimPair> prims;
    std::vector<double> pcoef;
    std::vector<double> schwarz;        // sqrt(max|(ij|ij)|) per pair id, filled on demand (device evaluation)
    double schwarz_thr = 1e-15;         // whole-job driver: skip work items whose quartets are all bounded below this (0 = off)
    // device tables
    PairHdr *d_pairs = nullptr;
    PrimPair *d_prims = nullptr;
    double *d_pcoef = nullptr, *d_rys = nullptr, *d_rys_fast = nullptr, *d_c2s = nullptr;
    // reusable work buffers
    Task *d_tasks = nullptr;        size_t cap_tasks = 0;
    double *d_out = nullptr;        size_t cap_out = 0;
    int *d_nonzero = nullptr;       size_t cap_nonzero = 0;
    double *d_scratch = nullptr;    size_t cap_scratch = 0;
    void *h_stage = nullptr;        size_t cap_stage = 0;
    void *d_ipwork[4] = {nullptr, nullptr, nullptr, nullptr};      // work tensors of the derivative blocks (deriv_block.cu), grow-only
    size_t cap_ipwork[4] = {0, 0, 0, 0};
    unsigned long long *d_counters = nullptr;
    long long launches = 0;
    struct JobPlan *plan = nullptr;     // cached whole-job plan (driver.cu)
    struct ListTables *ltab = nullptr;  // per-class pair tables of the list-mode fast path (driver.cu:list_mode_run)
    CINTOpt *deriv = nullptr;           // first-derivative helper context: shells [nbas, 2 nbas) = l+1 with coefficients -2 a c,
                                        // [2 nbas, 3 nbas) = l-1 (engine.cu:ctx_deriv), built on first use
    int profile = 0;                    // record per-launch events in the whole-job driver
    std::vector<double> profile_rows;
    int force_generic = 0;              // tests: route every class through the generic kernel
    int checksums = 0;                  // whole-job driver: reduce every tile to per-row sums (cintb200_set_checksums)
    std::mutex mtx;
    std::atomic<int> users{0};          // drop-in calls currently inside this context (cached contexts only; eviction waits for 0)
};

struct JobPlan;
void jobplan_free(JobPlan *p);
struct ListTables;
void listtables_free(ListTables *lt);
int list_mode_run(CINTOpt *c, const Task *tasks, size_t n, double *d_out, unsigned char *handled, int cart = 0);
int list_mode_device(CINTOpt *c, int ncenter, int cart, const int *shls, size_t n, const size_t *out_off, double *d_out_or_null,
                     double **d_out_used, size_t *total, int *nonzero);      // listdev.cu
int ctx_compute_schwarz(CINTOpt *c);
CINTOpt *ctx_deriv(CINTOpt *c);         // helper context with raised / lowered copies of every shell (engine.cu)
// dense shell-slice block on the tile kernels (driver.cu); ncenter 2, 3 or 4, cart: Cartesian output
int run_block(CINTOpt *c, int ncenter, const int *sl, double *out, int on_device, double *stats, int cart = 0);
int ctx_new_host(CINTOpt **out, const int *atm, int natm, const int *bas, int nbas, const double *env);
int b200_fail(int code, const char *fmt, ...);
// CINTB200_TIMING=1: host-phase timings on stderr (context build, plan build, job execution)
double b200_now();
bool b200_timing();
void b200_phase(const char *what, double t0);
int ctx_reserve(CINTOpt *c, void **ptr, size_t *cap, size_t bytes, bool pinned_host);
// large device buffers from the stream-ordered pool (kept across contexts; engine.cu)
int b200_big_alloc(void **p, size_t bytes);
void b200_big_free(void *p);
// Every device buffer of the library comes from that pool: a legacy cudaFree at context teardown was measured to stall for
// 0.1-0.9 s at random while 80 GB of pooled tiles are cached (bench.py's end-to-end step went from 1.6 to 2.2-2.4 s); pooled
// frees do not.  Same signatures as cudaMalloc / cudaFree.
template <class T> static inline cudaError_t b200_dmalloc(T **p, size_t bytes) { return b200_big_alloc((void **)p, bytes) ? cudaErrorMemoryAllocation : cudaSuccess; }
static inline cudaError_t b200_dfree(void *p) { b200_big_free(p); return cudaSuccess; }
