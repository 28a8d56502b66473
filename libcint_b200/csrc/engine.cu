// Host engine: context (= CINTOpt) construction, shell-pair tables, task building, class-sorted
// launches, and the C ABI of include/cint.h + include/cint_b200.h.
//
// Reference counterparts:  CINTall_2e_optimizer / CINTOpt_setij / CINTset_pairdata
// (src/optimizer.c:183,344,288), CINTinit_int2e_EnvVars (src/g2e.c:21), CINT2e_drv (src/cint2e.c:794),
// CINT3c2e_drv (src/cint3c2e.c:556), shell helpers (src/cint_bas.c), CINTgto_norm (src/misc.c:86).
// No CPU fallback exists: every entry point needs a CUDA device and fails loudly without one.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cstdarg>
#include <ctime>
#include <vector>
#include <map>
#include <mutex>
#include <algorithm>
#include <cuda_runtime.h>
#include "../../include/cint.h"
#include "../../include/cint_b200.h"
#include "types.h"
#include "kernels.h"
#include "rys.cuh"
#include "rys_tables.inc"
#include "rys_tables_fast.inc"
#include "c2s_tables.inc"
#include "engine.h"

static_assert(RYS_TAB_DEG == RYS_DEG && RYS_TAB_M == RYS_M && RYS_TAB_NMAX == RYS_NMAX, "rys.cuh out of sync with rys_tables.inc");
static_assert(C2S_LMAX >= B200_LMAX, "c2s table too small");
static_assert(RYS_FAST_DEG == RYS_FDEG && RYS_FAST_M == RYS_FM && RYS_FAST_NMAX == RYS_FNMAX && RYS_FAST_C0 == 8, "rys.cuh out of sync with rys_tables_fast.inc");

__constant__ RysMeta c_rys_meta;
__constant__ double c_rys_lx_r[RYS_NMAX * (RYS_NMAX + 1) / 2];
__constant__ double c_rys_lx_v[RYS_NMAX * (RYS_NMAX + 1) / 2];

static thread_local char g_err[512] = "";

int b200_fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    fprintf(stderr, "libcint_b200: error %d: %s\n", code, g_err);
    return code;
}
#define CUDA_OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return b200_fail(CINTB200_ENODEV, "%s failed: %s", #call, cudaGetErrorString(e_)); } while (0)

extern "C" const char *cintb200_last_error(void) { return g_err; }

double b200_now()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
bool b200_timing() { static const bool on = getenv("CINTB200_TIMING") && atoi(getenv("CINTB200_TIMING")); return on; }
void b200_phase(const char *what, double t0) { if (b200_timing()) fprintf(stderr, "[cintb200 timing] %-28s %8.2f ms\n", what, 1e3 * (b200_now() - t0)); }

// ------------------------------------------------------------------ shell helpers (src/cint_bas.c)
extern "C" {
FINT CINTlen_cart(const FINT l) { return (l + 1) * (l + 2) / 2; }
FINT CINTcgtos_cart(const FINT b, const FINT *bas) { return CINTlen_cart(bas(ANG_OF, b)) * bas(NCTR_OF, b); }
FINT CINTcgto_cart(const FINT b, const FINT *bas) { return CINTcgtos_cart(b, bas); }
FINT CINTcgtos_spheric(const FINT b, const FINT *bas) { return (bas(ANG_OF, b) * 2 + 1) * bas(NCTR_OF, b); }
FINT CINTcgto_spheric(const FINT b, const FINT *bas) { return CINTcgtos_spheric(b, bas); }
FINT CINTtot_pgto_spheric(const FINT *bas, const FINT nbas)
{
    FINT s = 0;
    for (FINT i = 0; i < nbas; i++) s += (bas(ANG_OF, i) * 2 + 1) * bas(NPRIM_OF, i);
    return s;
}
FINT CINTtot_cgto_spheric(const FINT *bas, const FINT nbas)
{
    FINT s = 0;
    for (FINT i = 0; i < nbas; i++) s += CINTcgtos_spheric(i, bas);
    return s;
}
FINT CINTtot_cgto_cart(const FINT *bas, const FINT nbas)
{
    FINT s = 0;
    for (FINT i = 0; i < nbas; i++) s += CINTcgtos_cart(i, bas);
    return s;
}
void CINTshells_cart_offset(FINT ao_loc[], const FINT *bas, const FINT nbas)
{
    ao_loc[0] = 0;
    for (FINT i = 1; i < nbas; i++) ao_loc[i] = ao_loc[i - 1] + CINTcgtos_cart(i - 1, bas);
}
void CINTshells_spheric_offset(FINT ao_loc[], const FINT *bas, const FINT nbas)
{
    ao_loc[0] = 0;
    for (FINT i = 1; i < nbas; i++) ao_loc[i] = ao_loc[i - 1] + CINTcgtos_spheric(i - 1, bas);
}
// 1/sqrt( int_0^inf r^(2l+2) exp(-2 a r^2) dr ) = 1/sqrt( Gamma(l+3/2) / (2 (2a)^(l+3/2)) )
double CINTgto_norm(FINT n, double a)
{
    double p = n + 1.5;
    return 1.0 / sqrt(tgamma(p) / (2.0 * pow(2.0 * a, p)));
}
}

// ------------------------------------------------------------------ context
static uint64_t fnv1a(const void *data, size_t n, uint64_t h)
{
    const unsigned char *p = (const unsigned char *)data;
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ULL; }
    return h;
}

static size_t env_extent(const int *atm, int natm, const int *bas, int nbas)
{
    size_t n = PTR_ENV_START;
    for (int i = 0; i < natm; i++) n = std::max(n, (size_t)atm(PTR_COORD, i) + 3);
    for (int i = 0; i < nbas; i++) {
        n = std::max(n, (size_t)bas(PTR_EXP, i) + bas(NPRIM_OF, i));
        n = std::max(n, (size_t)bas(PTR_COEFF, i) + (size_t)bas(NPRIM_OF, i) * bas(NCTR_OF, i));
    }
    return n;
}

static uint64_t basis_hash(const int *atm, int natm, const int *bas, int nbas, const double *env)
{
    uint64_t h = 1469598103934665603ULL;
    h = fnv1a(&natm, sizeof natm, h);
    h = fnv1a(&nbas, sizeof nbas, h);
    for (int i = 0; i < natm; i++) h = fnv1a(&atm(PTR_COORD, i), sizeof(int), h);
    for (int i = 0; i < nbas; i++) {
        int s[6] = {bas(ATOM_OF, i), bas(ANG_OF, i), bas(NPRIM_OF, i), bas(NCTR_OF, i), bas(PTR_EXP, i), bas(PTR_COEFF, i)};
        h = fnv1a(s, sizeof s, h);
    }
    h = fnv1a(env, sizeof(double) * env_extent(atm, natm, bas, nbas), h);
    return h;
}

static int g_constants_ready_dev[64];

static std::mutex g_constants_mtx;

static int setup_device_constants(int dev)
{
    std::lock_guard<std::mutex> lock(g_constants_mtx);       // contexts may be created from several threads at once
    if (dev < 64 && g_constants_ready_dev[dev]) return 0;
    RysMeta meta;
    memset(&meta, 0, sizeof meta);
    for (int n = 1; n <= RYS_NMAX; n++) { meta.off[n] = RYS_TAB_OFF[n]; meta.nint[n] = RYS_TAB_NINT[n]; }
    for (int n = 1; n <= RYS_FNMAX; n++) meta.fast_nint[n] = RYS_FAST_NINT[n];
    CUDA_OK(cudaMemcpyToSymbol(c_rys_meta, &meta, sizeof meta));
    CUDA_OK(cudaMemcpyToSymbol(c_rys_lx_r, RYS_LX_R, sizeof(double) * RYS_NMAX * (RYS_NMAX + 1) / 2));
    CUDA_OK(cudaMemcpyToSymbol(c_rys_lx_v, RYS_LX_V, sizeof(double) * RYS_NMAX * (RYS_NMAX + 1) / 2));
    if (generic_setup_constants()) return b200_fail(CINTB200_ENODEV, "constant upload failed");
    if (dev < 64) g_constants_ready_dev[dev] = 1;
    return 0;
}

static void build_pairs(CINTOpt *c)
{
    const int nbas = c->nbas;
    const int *bas = c->bas.data();
    const int *atm = c->atm.data();
    const double *env = c->env.data();
    c->shells.resize(nbas);
    std::vector<std::vector<double>> logmaxc(nbas);
    int ao_s = 0, ao_c = 0;
    for (int i = 0; i < nbas; i++) {
        ShellInfo &s = c->shells[i];
        s.l = bas(ANG_OF, i); s.nprim = bas(NPRIM_OF, i); s.nctr = bas(NCTR_OF, i);
        s.r = env + atm(PTR_COORD, bas(ATOM_OF, i));
        s.exps = env + bas(PTR_EXP, i);
        s.coef = env + bas(PTR_COEFF, i);
        s.ao_sph = ao_s; s.ao_cart = ao_c;
        ao_s += (2 * s.l + 1) * s.nctr;
        ao_c += B200_NCART(s.l) * s.nctr;
        logmaxc[i].resize(s.nprim);
        for (int p = 0; p < s.nprim; p++) {             // CINTOpt_log_max_pgto_coeff, src/optimizer.c:259
            double mx = 0;
            for (int k = 0; k < s.nctr; k++) mx = std::max(mx, fabs(s.coef[k * s.nprim + p]));
            logmaxc[i][p] = log(mx);
        }
    }
    c->nao_sph = ao_s; c->nao_cart = ao_c;
    const size_t npair2 = (size_t)nbas * (nbas + 1) / 2;
    c->pairs.assign(npair2 + nbas, PairHdr());
    c->prims.clear();
    c->pcoef.clear();
    const double omega = c->omega;
    for (int i = 0; i < nbas; i++)
        for (int j = 0; j <= i; j++) {
            PairHdr &h = c->pairs[(size_t)i * (i + 1) / 2 + j];
            int a = i, b = j;
            if (c->shells[j].l > c->shells[i].l) { a = j; b = i; }
            const ShellInfo &sa = c->shells[a], &sb = c->shells[b];
            h.sh_a = a; h.sh_b = b; h.la = sa.l; h.lb = sb.l; h.nca = sa.nctr; h.ncb = sb.nctr;
            h.ao_a = sa.ao_sph; h.ao_b = sb.ao_sph;
            h.pp_off = (int)c->prims.size();
            h.cc_off = (int)c->pcoef.size();
            double rr = 0;
            for (int d = 0; d < 3; d++) { h.ra[d] = sa.r[d]; h.ab[d] = sa.r[d] - sb.r[d]; rr += h.ab[d] * h.ab[d]; }
            // CINTset_pairdata, src/optimizer.c:301-314
            double log_rr = 1.7 - 1.5 * log(sa.exps[sa.nprim - 1] + sb.exps[sb.nprim - 1]);
            const int lij = sa.l + sb.l;
            if (lij > 0) {
                double dist = sqrt(rr);
                if (omega < 0) {
                    double th = omega * omega / (omega * omega + sa.exps[sa.nprim - 1] + sb.exps[sb.nprim - 1]);
                    log_rr += lij * log(dist + th * 8. + 1.);
                } else {
                    log_rr += lij * log(dist + 1.);
                }
            }
            int npp = 0;
            for (int jp = 0; jp < sb.nprim; jp++)
                for (int ip = 0; ip < sa.nprim; ip++) {
                    const double aa = sa.exps[ip], ab = sb.exps[jp];
                    const double aij = aa + ab;
                    const double eij = rr * aa * ab / aij;
                    const double cce = eij - log_rr - logmaxc[a][ip] - logmaxc[b][jp];
                    if (!(cce < c->expcutoff4)) continue;
                    PrimPair pp;
                    pp.aij = aij;
                    const double wj = ab / aij;
                    pp.px = sa.r[0] - wj * h.ab[0];
                    pp.py = sa.r[1] - wj * h.ab[1];
                    pp.pz = sa.r[2] - wj * h.ab[2];
                    pp.kij = exp(-eij);
                    pp.cce = cce;
                    pp.inv_aij = 1.0 / aij;
                    pp.ipa = ip; pp.ipb = jp;
                    c->prims.push_back(pp);
                    for (int cb = 0; cb < sb.nctr; cb++)
                        for (int ca = 0; ca < sa.nctr; ca++)
                            c->pcoef.push_back(sa.coef[ca * sa.nprim + ip] * sb.coef[cb * sb.nprim + jp]);
                    npp++;
                }
            h.npp = npp;
        }
    // single-shell pseudo pairs: ket of (ij|k).  al = 0, rkl = rk, ekl = 1 (src/g3c2e.c:96-105);
    // kij carries 1/fac_sp(0) = 2 sqrt(pi) so kernels can apply fac_sp(0) for the absent shell uniformly.
    for (int k = 0; k < nbas; k++) {
        PairHdr &h = c->pairs[npair2 + k];
        const ShellInfo &s = c->shells[k];
        h.sh_a = k; h.sh_b = -1; h.la = s.l; h.lb = 0; h.nca = s.nctr; h.ncb = 1;
        h.ao_a = s.ao_sph; h.ao_b = 0;
        h.pp_off = (int)c->prims.size();
        h.cc_off = (int)c->pcoef.size();
        for (int d = 0; d < 3; d++) { h.ra[d] = s.r[d]; h.ab[d] = 0; }
        for (int p = 0; p < s.nprim; p++) {
            PrimPair pp;
            pp.aij = s.exps[p];
            pp.px = s.r[0]; pp.py = s.r[1]; pp.pz = s.r[2];
            pp.kij = 3.5449077018110320546;
            pp.cce = 0;
            pp.inv_aij = 1.0 / pp.aij;
            pp.ipa = p; pp.ipb = 0;
            c->prims.push_back(pp);
            for (int ca = 0; ca < s.nctr; ca++) c->pcoef.push_back(s.coef[ca * s.nprim + p]);
        }
        h.npp = s.nprim;
    }
    // Virtual segmented pairs.  A generally contracted pair whose type (la, lb, nca x ncb) has no specialised kernel in any
    // class -- general contractions above s, e.g. cc-pVXZ of second-row atoms, ANO sets, the raised copies of contracted
    // shells in the derivative context -- would send every class it takes part in to the catch-all kernel.  Such a pair also
    // gets nca x ncb headers with ONE contraction each (same primitives, own coefficient column): the dense-block driver
    // (driver.cu:run_block) addresses them as separate sub-blocks, which recomputes the primitives per contraction but on the
    // register / cooperative kernels (~100x faster per primitive quartet than the catch-all kernel).
    c->vfirst.assign(npair2, -1);
    for (size_t p = 0; p < npair2; p++) {
        const PairHdr h = c->pairs[p];                  // copy: the vector grows below
        const int ncomb = h.nca * h.ncb;
        if (ncomb <= 1 || h.npp == 0) continue;
        CoopInfo ci;
        auto specialised = [&](int nc) { return reg_kernel_lookup(h.la, h.lb, 0, 0, nc, 1) != nullptr || coop_kernel_lookup(h.la, h.lb, 0, 0, nc, 1, &ci) != nullptr; };
        if (specialised(ncomb) || !specialised(1)) continue;       // contracted type covered, or not even the segmented type is
        c->vfirst[p] = (int)c->pairs.size();
        for (int cb = 0; cb < h.ncb; cb++)
            for (int ca = 0; ca < h.nca; ca++) {
                PairHdr v = h;
                v.nca = v.ncb = 1;
                v.cc_off = (int)c->pcoef.size();
                for (int q = 0; q < h.npp; q++) c->pcoef.push_back(c->pcoef[(size_t)h.cc_off + (size_t)q * ncomb + cb * h.nca + ca]);
                c->pairs.push_back(v);
            }
    }
}

static int ctx_upload(CINTOpt *c)
{
    CUDA_OK(b200_dmalloc(&c->d_pairs, sizeof(PairHdr) * c->pairs.size()));
    CUDA_OK(b200_dmalloc(&c->d_prims, sizeof(PrimPair) * std::max<size_t>(1, c->prims.size())));
    CUDA_OK(b200_dmalloc(&c->d_pcoef, sizeof(double) * std::max<size_t>(1, c->pcoef.size())));
    CUDA_OK(b200_dmalloc(&c->d_rys, sizeof(RYS_TAB_COEF)));
    CUDA_OK(b200_dmalloc(&c->d_rys_fast, sizeof(RYS_FAST_COEF)));
    CUDA_OK(cudaMemcpy(c->d_rys_fast, RYS_FAST_COEF, sizeof(RYS_FAST_COEF), cudaMemcpyHostToDevice));
    CUDA_OK(b200_dmalloc(&c->d_c2s, sizeof(C2S_COEF)));
    CUDA_OK(cudaMemcpy(c->d_pairs, c->pairs.data(), sizeof(PairHdr) * c->pairs.size(), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(c->d_prims, c->prims.data(), sizeof(PrimPair) * c->prims.size(), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(c->d_pcoef, c->pcoef.data(), sizeof(double) * c->pcoef.size(), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(c->d_rys, RYS_TAB_COEF, sizeof(RYS_TAB_COEF), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(c->d_c2s, C2S_COEF, sizeof(C2S_COEF), cudaMemcpyHostToDevice));
    CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    return 0;
}

// host-only part of context construction (validation, copies of the arrays, shell-pair tables)
int ctx_new_host(CINTOpt **out, const int *atm, int natm, const int *bas, int nbas, const double *env)
{
    if (!out || !atm || !bas || !env || natm <= 0 || nbas <= 0) return b200_fail(CINTB200_EINVAL, "cintb200_create: bad arguments");
    *out = NULL;
    for (int i = 0; i < nbas; i++) {
        if (bas(ANG_OF, i) < 0 || bas(ANG_OF, i) > B200_LMAX)
            return b200_fail(CINTB200_EINVAL, "shell %d: angular momentum %d outside 0..%d", i, bas(ANG_OF, i), B200_LMAX);
        if (bas(NPRIM_OF, i) < 1 || bas(NCTR_OF, i) < 1 || bas(ATOM_OF, i) < 0 || bas(ATOM_OF, i) >= natm)
            return b200_fail(CINTB200_EINVAL, "shell %d: malformed bas entry", i);
    }
    CINTOpt *c = new CINTOpt();
    c->magic = B200_CTX_MAGIC;
    c->device = -1;
    c->natm = natm; c->nbas = nbas;
    c->atm.assign(atm, atm + (size_t)natm * ATM_SLOTS);
    c->bas.assign(bas, bas + (size_t)nbas * BAS_SLOTS);
    c->env.assign(env, env + env_extent(atm, natm, bas, nbas));
    c->hash = basis_hash(atm, natm, bas, nbas, env);
    // CINTinit_int2e_EnvVars src/g2e.c:57-62 (4c: +1 when user-set), CINTinit_int3c2e_EnvVars src/g3c2e.c:55-59
    const double e0 = env[PTR_EXPCUTOFF];
    c->expcutoff4 = (e0 == 0) ? 60.0 : std::max(40.0, e0) + 1.0;
    c->expcutoff3 = (e0 == 0) ? 60.0 : std::max(40.0, e0);
    c->omega = env[PTR_RANGE_OMEGA];
    build_pairs(c);
    *out = c;
    return 0;
}

extern "C" int cintb200_create(cintb200_ctx **out, const int *atm, int natm, const int *bas, int nbas,
                               const double *env, int device)
{
    if (!out) return b200_fail(CINTB200_EINVAL, "cintb200_create: bad arguments");
    *out = NULL;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return b200_fail(CINTB200_ENODEV, "no CUDA device available (%s); this library has no CPU path",
                         e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    int caller_dev = -1;
    CUDA_OK(cudaGetDevice(&caller_dev));
    if (device < 0) device = caller_dev;
    if (device >= ndev) return b200_fail(CINTB200_EINVAL, "device %d out of range (%d devices)", device, ndev);
    struct RestoreDevice { int d; ~RestoreDevice() { if (d >= 0) cudaSetDevice(d); } } restore{device != caller_dev ? caller_dev : -1};
    CUDA_OK(cudaSetDevice(device));
    if (setup_device_constants(device)) return CINTB200_ENODEV;
    CINTOpt *c = NULL;
    double t0 = b200_now();
    int rc = ctx_new_host(&c, atm, natm, bas, nbas, env);
    if (rc) return rc;
    b200_phase("context: pair tables (host)", t0);
    c->device = device;
    t0 = b200_now();
    rc = ctx_upload(c);
    b200_phase("context: upload", t0);
    if (rc) { cintb200_destroy(c); return rc; }
    *out = c;
    return 0;
}

extern "C" void cintb200_destroy(cintb200_ctx *c)
{
    if (!c || c->magic != B200_CTX_MAGIC) return;
    if (c->device < 0) {            // host-only context (planning without a GPU)
        if (c->plan) { jobplan_free(c->plan); c->plan = nullptr; }
        c->magic = 0;
        delete c;
        return;
    }
    cudaSetDevice(c->device);
    double tph = b200_now();
    if (c->stream) cudaStreamSynchronize(c->stream);
    b200_phase("destroy: stream sync", tph);
    tph = b200_now();
    b200_dfree(c->d_pairs); b200_dfree(c->d_prims); b200_dfree(c->d_pcoef); b200_dfree(c->d_rys); b200_dfree(c->d_rys_fast); b200_dfree(c->d_c2s);
    b200_dfree(c->d_tasks); b200_dfree(c->d_out); b200_dfree(c->d_nonzero); b200_dfree(c->d_scratch); b200_dfree(c->d_counters);
    b200_phase("destroy: context tables", tph);
    tph = b200_now();
    if (c->plan) { jobplan_free(c->plan); c->plan = nullptr; }
    b200_phase("destroy: job plan", tph);
    if (c->deriv) { cintb200_destroy(c->deriv); c->deriv = nullptr; }
    if (c->ltab) { listtables_free(c->ltab); c->ltab = nullptr; }
    if (c->h_stage) cudaFreeHost(c->h_stage);
    for (int k = 0; k < 4; k++) b200_dfree(c->d_ipwork[k]);
    if (c->stream) cudaStreamDestroy(c->stream);
    c->magic = 0;
    delete c;
}

extern "C" int cintb200_device(const cintb200_ctx *c) { return (c && c->magic == B200_CTX_MAGIC) ? c->device : -1; }

int ctx_reserve(CINTOpt *c, void **ptr, size_t *cap, size_t bytes, bool pinned_host)
{
    if (*cap >= bytes) return 0;
    size_t want = std::max(bytes, *cap * 2);
    if (*ptr) { if (pinned_host) cudaFreeHost(*ptr); else b200_dfree(*ptr); *ptr = NULL; *cap = 0; }
    cudaError_t e = pinned_host ? cudaMallocHost(ptr, want) : b200_dmalloc(ptr, want);
    if (e != cudaSuccess && want > bytes) { want = bytes; e = pinned_host ? cudaMallocHost(ptr, want) : b200_dmalloc(ptr, want); }
    if (e != cudaSuccess) return b200_fail(CINTB200_ENOMEM, "allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
    *cap = want;
    (void)c;
    return 0;
}

// ------------------------------------------------------------------ list-mode batches
struct ClassKey {
    int la, lb, lc, ld, ncab, nccd;
    bool operator<(const ClassKey &o) const
    { return memcmp(this, &o, sizeof(ClassKey)) < 0; }
};

static inline int shell_dim(const ShellInfo &s, int cart) { return (cart ? B200_NCART(s.l) : 2 * s.l + 1) * s.nctr; }
static long run_batch(CINTOpt *c, int ncenter, int kind, const int *shls, size_t n, const size_t *out_off,
                      double *out, int on_device, int *nonzero, int cart_pos = -1);

extern "C" size_t cintb200_block_size(const cintb200_ctx *c, int kind, const int *shls, int ncenter)
{
    if (!c || c->magic != B200_CTX_MAGIC) return 0;
    size_t n = 1;
    for (int m = 0; m < ncenter; m++) {
        if (shls[m] < 0 || shls[m] >= c->nbas) return 0;
        n *= shell_dim(c->shells[shls[m]], kind == CINTB200_CART);
    }
    return n;
}

// cart_pos >= 0: the shell at that position of every tuple keeps Cartesian components while the others are transformed
// to the requested kind (used by the first-derivative assembly, which differentiates in the Cartesian basis)
static long run_batch(CINTOpt *c, int ncenter, int kind, const int *shls, size_t n, const size_t *out_off,
                      double *out, int on_device, int *nonzero, int cart_pos)
{
    if (!c || c->magic != B200_CTX_MAGIC) return b200_fail(CINTB200_EINVAL, "invalid context");
    if (n == 0) return 0;
    if (!shls || !out) return b200_fail(CINTB200_EINVAL, "NULL shls/out");
    const int cart = (kind == CINTB200_CART);
    std::lock_guard<std::mutex> lock(c->mtx);
    CUDA_OK(cudaSetDevice(c->device));
    const size_t npair2 = (size_t)c->nbas * (c->nbas + 1) / 2;

    // Large lists: the whole bookkeeping runs on the device (listdev.cu) when every class of the list has a specialised kernel
    static const bool list_fast = getenv("CINTB200_LIST_GENERIC") == nullptr;
    static const bool list_dev = getenv("CINTB200_LIST_HOST") == nullptr;
    if (list_fast && list_dev && n >= 4096 && cart_pos < 0 && !c->force_generic) {
        size_t tot_dev = 0;
        double *d_used = nullptr;
        const int r = list_mode_device(c, ncenter, cart, shls, n, out_off, on_device ? out : nullptr, &d_used, &tot_dev, nonzero);
        if (r < 0) return r;
        if (r == 1) {
            if (!on_device) {
                if (ctx_reserve(c, (void **)&c->h_stage, &c->cap_stage, sizeof(double) * std::max<size_t>(1, tot_dev), true)) return CINTB200_ENOMEM;
                CUDA_OK(cudaMemcpyAsync(c->h_stage, d_used, sizeof(double) * tot_dev, cudaMemcpyDeviceToHost, c->stream));
            }
            CUDA_OK(cudaStreamSynchronize(c->stream));
            cudaError_t le = cudaGetLastError();
            if (le != cudaSuccess) return b200_fail(CINTB200_ENODEV, "kernel execution failed: %s", cudaGetErrorString(le));
            if (!on_device) {
                if (out_off) {
                    for (size_t t = 0; t < n; t++) {
                        const size_t len = cintb200_block_size(c, kind, shls + t * ncenter, ncenter);
                        memcpy(out + out_off[t], (double *)c->h_stage + out_off[t], sizeof(double) * len);
                    }
                } else memcpy(out, c->h_stage, sizeof(double) * tot_dev);
            }
            return (long)n;
        }
    }
    std::vector<Task> tasks(n);
    std::vector<ClassKey> keys(n);
    std::vector<size_t> offs(n);
    size_t total = 0;
    // pass 1 (host threads): validation and block sizes; offsets are a prefix sum unless the caller gave them
    {
        long long bad = -1;
#pragma omp parallel for schedule(static) if (n > 20000)
        for (long long t = 0; t < (long long)n; t++) {
            const int *s = shls + t * ncenter;
            size_t sz = 1;
            for (int m = 0; m < ncenter; m++) {
                if (s[m] < 0 || s[m] >= c->nbas) {
#pragma omp critical
                    bad = (bad < 0 || t < bad) ? t : bad;
                    sz = 0;
                    break;
                }
                sz *= shell_dim(c->shells[s[m]], cart || cart_pos == m);
            }
            offs[t] = sz;
        }
        if (bad >= 0) return b200_fail(CINTB200_EINVAL, "tuple %lld: shell id out of range", bad);
        for (size_t t = 0; t < n; t++) {
            const size_t sz = offs[t];
            offs[t] = out_off ? out_off[t] : total;
            total = std::max(total, offs[t] + sz);
        }
    }
    // pass 2 (host threads): pair ids, strides, class keys
#pragma omp parallel for schedule(static) if (n > 20000)
    for (long long t = 0; t < (long long)n; t++) {
        const int *s = shls + t * ncenter;
        Task &T = tasks[t];
        if (ncenter == 2) {
            // (i|k): both sides are single-shell pseudo pairs (aj = al = 0, src/g2c2e.c:15-100); block (di,dk)
            const int i = s[0], k = s[1];
            const long long di = shell_dim(c->shells[i], cart || cart_pos == 0), dk = shell_dim(c->shells[k], cart || cart_pos == 1);
            T.bra = (int)(npair2 + i); T.ket = (int)(npair2 + k);
            T.sa = 1; T.sb = 0; T.sc = di; T.sd = 0;
            if (cart_pos >= 0) T.flags = cart_pos == 0 ? 1 : 4;
            T.off = (long long)offs[t];
            const PairHdr &hb = c->pairs[T.bra], &hk = c->pairs[T.ket];
            keys[t] = ClassKey{hb.la, hb.lb, hk.la, hk.lb, hb.nca * hb.ncb, hk.nca * hk.ncb};
            continue;
        }
        const int i = s[0], j = s[1], k = s[2], l = (ncenter == 4) ? s[3] : -1;
        const long long di = shell_dim(c->shells[i], cart || cart_pos == 0), dj = shell_dim(c->shells[j], cart || cart_pos == 1);
        const long long dk = shell_dim(c->shells[k], cart || cart_pos == 2), dl = (l >= 0) ? shell_dim(c->shells[l], cart || cart_pos == 3) : 1;
        T.bra = (int)((i >= j) ? (size_t)i * (i + 1) / 2 + j : (size_t)j * (j + 1) / 2 + i);
        const PairHdr &hb = c->pairs[T.bra];
        const long long si = 1, sj = di, sk = di * dj, sl = di * dj * dk;
        if (hb.sh_a == i && (i != j || true)) { T.sa = (int)si; T.sb = (int)sj; }
        if (hb.sh_a != i) { T.sa = (int)sj; T.sb = (int)si; }
        if (cart_pos == 0) T.flags = (hb.sh_a == i) ? 1 : 2;
        if (cart_pos == 1) T.flags = (hb.sh_a == j && hb.sh_a != i) ? 1 : 2;
        if (l >= 0) {
            T.ket = (int)((k >= l) ? (size_t)k * (k + 1) / 2 + l : (size_t)l * (l + 1) / 2 + k);
            const PairHdr &hk = c->pairs[T.ket];
            if (hk.sh_a == k) { T.sc = sk; T.sd = sl; } else { T.sc = sl; T.sd = sk; }
            if (cart_pos == 2) T.flags = (hk.sh_a == k) ? 4 : 8;
            if (cart_pos == 3) T.flags = (hk.sh_a == l && hk.sh_a != k) ? 4 : 8;
        } else {
            T.ket = (int)(npair2 + k);
            T.sc = sk; T.sd = 0;
            if (cart_pos == 2) T.flags = 4;
        }
        const PairHdr &hk = c->pairs[T.ket];
        T.off = (long long)offs[t];
        (void)dl;
        keys[t] = ClassKey{hb.la, hb.lb, hk.la, hk.lb, hb.nca * hb.ncb, hk.nca * hk.ncb};
    }
    double *d_out = out;
    if (!on_device) {
        if (ctx_reserve(c, (void **)&c->d_out, &c->cap_out, sizeof(double) * total, false)) return CINTB200_ENOMEM;
        d_out = c->d_out;
    }
    // fast path: tuples whose classes have a specialised tile kernel run there (driver.cu:list_mode_run);
    // packed output below 2^31 elements (the kernels keep row offsets in 32 bits)
    std::vector<unsigned char> handled(n, 0);
    if (list_fast && n >= 32 && cart_pos < 0 && !c->force_generic && total < ((size_t)1 << 31)) {   // single calls: generic kernel (lower latency)
        if (list_mode_run(c, tasks.data(), n, d_out, handled.data(), cart)) return CINTB200_ENODEV;
    }
    // the rest: class-sorted order for the generic kernel
    const size_t n_all = n;
    std::vector<size_t> order;
    order.reserve(n);
    for (size_t t = 0; t < n; t++) if (!handled[t]) order.push_back(t);
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return keys[a] < keys[b]; });
    n = order.size();
    std::vector<Task> sorted(n);
    for (size_t t = 0; t < n; t++) sorted[t] = tasks[order[t]];

    if (ctx_reserve(c, (void **)&c->d_tasks, &c->cap_tasks, sizeof(Task) * std::max<size_t>(1, n), false)) return CINTB200_ENOMEM;
    if (ctx_reserve(c, (void **)&c->d_nonzero, &c->cap_nonzero, sizeof(int) * std::max<size_t>(1, n), false)) return CINTB200_ENOMEM;
    if (n) CUDA_OK(cudaMemcpyAsync(c->d_tasks, sorted.data(), sizeof(Task) * n, cudaMemcpyHostToDevice, c->stream));

    EngineParams P;
    P.pairs = c->d_pairs; P.prims = c->d_prims; P.pcoef = c->d_pcoef; P.rys_coef = c->d_rys; P.c2s = c->d_c2s;
    P.expcutoff = (ncenter == 4) ? c->expcutoff4 : c->expcutoff3;
    P.omega = c->omega;
    P.cart = cart;

    size_t start = 0;
    while (start < n) {
        size_t end = start + 1;
        const ClassKey &k0 = keys[order[start]];
        while (end < n && !(k0 < keys[order[end]]) && !(keys[order[end]] < k0)) end++;
        GenericClass C;
        GenericLaunch L;
        if (generic_plan(&C, &L, k0.la, k0.lb, k0.lc, k0.ld, k0.ncab, k0.nccd, cart, (long long)(end - start), C2S_OFF, c->omega < 0))
            return b200_fail(CINTB200_ENOSUP, "class (%d%d|%d%d) exceeds this build's limits (nroots <= %d)",
                             k0.la, k0.lb, k0.lc, k0.ld, RYS_NMAX);
        if (C.scratch_per_block) {
            if (ctx_reserve(c, (void **)&c->d_scratch, &c->cap_scratch, sizeof(double) * C.scratch_per_block * L.grid, false))
                return CINTB200_ENOMEM;
            C.scratch = c->d_scratch;
        }
        if (generic_launch(P, C, L, c->d_tasks + start, (long long)(end - start), d_out, c->d_nonzero + start, NULL, c->stream))
            return b200_fail(CINTB200_ENODEV, "kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        c->launches++;
        if (C.scratch_per_block) CUDA_OK(cudaStreamSynchronize(c->stream));   // scratch may be regrown by the next class
        start = end;
    }
    if (!on_device) {
        // device -> pinned staging -> caller's pageable memory
        if (ctx_reserve(c, (void **)&c->h_stage, &c->cap_stage, sizeof(double) * total, true)) return CINTB200_ENOMEM;
        CUDA_OK(cudaMemcpyAsync(c->h_stage, d_out, sizeof(double) * total, cudaMemcpyDeviceToHost, c->stream));
    }
    std::vector<int> nz;
    if (nonzero && n) {
        nz.resize(n);
        CUDA_OK(cudaMemcpyAsync(nz.data(), c->d_nonzero, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
    }
    CUDA_OK(cudaStreamSynchronize(c->stream));
    {
        cudaError_t le = cudaGetLastError();
        if (le != cudaSuccess) return b200_fail(CINTB200_ENODEV, "kernel execution failed: %s", cudaGetErrorString(le));
    }
    if (!on_device) {
        if (out_off) {
            for (size_t t = 0; t < n_all; t++) {
                size_t len = cintb200_block_size(c, kind, shls + t * ncenter, ncenter);
                if (cart_pos >= 0) len = len / shell_dim(c->shells[shls[t * ncenter + cart_pos]], cart) * shell_dim(c->shells[shls[t * ncenter + cart_pos]], 1);
                memcpy(out + offs[t], (double *)c->h_stage + offs[t], sizeof(double) * len);
            }
        } else {
            memcpy(out, c->h_stage, sizeof(double) * total);
        }
    }
    if (nonzero) {
        // tile kernels screen at the pair level only: a block is non-empty when both pairs kept a primitive
        for (size_t t = 0; t < n_all; t++) if (handled[t]) nonzero[t] = c->pairs[tasks[t].bra].npp > 0 && c->pairs[tasks[t].ket].npp > 0;
        for (size_t t = 0; t < n; t++) nonzero[order[t]] = nz[t];
    }
    return (long)n_all;
}

extern "C" long cintb200_int2e_batch(cintb200_ctx *c, int kind, const int *shls, size_t n, const size_t *out_off,
                                     double *out, int on_device, int *nonzero)
{ return run_batch(c, 4, kind, shls, n, out_off, out, on_device, nonzero); }

extern "C" long cintb200_int3c2e_batch(cintb200_ctx *c, int kind, const int *shls, size_t n, const size_t *out_off,
                                       double *out, int on_device, int *nonzero)
{ return run_batch(c, 3, kind, shls, n, out_off, out, on_device, nonzero); }

extern "C" long cintb200_int2c2e_batch(cintb200_ctx *c, int kind, const int *shls, size_t n, const size_t *out_off,
                                       double *out, int on_device, int *nonzero)
{ return run_batch(c, 2, kind, shls, n, out_off, out, on_device, nonzero); }

// ------------------------------------------------------------------ first derivatives
// int2e_ip1 / int3c2e_ip1: ( nabla i j | k l ), src/autocode/grad2.c:19-68 and src/autocode/int3c2e.c (ng = {1,0,0,0,1,1,1,3}:
// one extra unit of angular momentum on i, 3 tensor components).  The reference differentiates on the g array,
//     d/dx [ (x-X)^n e^{-a (x-X)^2} ] = n (x-X)^{n-1} e^{..} - 2 a (x-X)^{n+1} e^{..}          (CINTnabla1i_2e, src/g2e.c:4550)
// Here the same identity is applied one level up, on whole SHELLS: a helper context holds for every shell i a raised shell
// i+ (l+1, contraction coefficients -2 a_p c_kp) and a lowered shell i- (l-1, same coefficients).  The Cartesian blocks
// (i+ j|kl) and (i- j|kl) come from the ordinary engine (first index left in Cartesians), and one small kernel combines
// them component by component, applies cart->sph on i and writes the three blocks out[comp][l][k][j][i].
struct IpTask {
    size_t off_p, off_m, off_o, comp_stride;   // element offsets of the raised / lowered / output blocks, stride between components
    int li, nctr, rest, has_m;                 // rest = product of the dimensions AFTER the differentiated index (contraction included)
    int post, pad;                             // product of the dimensions BEFORE it (the faster-running indices)
};

__device__ __forceinline__ void ip_cart_xyz(int l, int idx, int &lx, int &ly, int &lz)
{
    int n = 0;
    for (lx = l; lx >= 0; lx--) {
        const int cnt = l - lx + 1;
        if (idx < n + cnt) { lz = idx - n; ly = l - lx - lz; return; }
        n += cnt;
    }
    lx = ly = lz = 0;
}
__device__ __forceinline__ int ip_cart_index(int lx, int lz, int l) { const int r = l - lx; return r * (r + 1) / 2 + lz; }

__global__ void ip1_assemble_kernel(const IpTask *__restrict__ tasks, size_t ntasks, const double *__restrict__ bp,
                                    const double *__restrict__ bm, double *__restrict__ out, const double *__restrict__ c2s,
                                    const int *__restrict__ c2s_off, int cart)
{
    const double fsp[2] = {0.282094791773878143, 0.488602511902919921};
    for (size_t t = blockIdx.x; t < ntasks; t += gridDim.x) {
        const IpTask T = tasks[t];
        const int li = T.li, nfi = B200_NCART(li), nfp = B200_NCART(li + 1), nfm = li > 0 ? B200_NCART(li - 1) : 0;
        const bool sph = !cart && li >= 2;
        const int di = sph ? 2 * li + 1 : nfi;
        // the engine scales s and p functions by fac_sp(l) (src/g1e.c:565-572) instead of transforming them: undo it for
        // the raised / lowered shell and apply the factor of the target shell
        const double fi = li < 2 ? fsp[li] : 1.0;
        const double sp = fi / (li + 1 < 2 ? fsp[li + 1] : 1.0);
        const double sm = li > 0 ? fi / (li - 1 < 2 ? fsp[li - 1] : 1.0) : 0.0;
        const double *cm = c2s + c2s_off[li];
        const size_t post = (size_t)T.post;
        const size_t per = post * di * T.nctr * T.rest;
        for (size_t idx = threadIdx.x; idx < 3 * per; idx += blockDim.x) {
            const int comp = (int)(idx / per);
            size_t w = idx - (size_t)comp * per;
            const size_t q = w % post;
            w /= post;
            const int m = (int)(w % di);
            w /= di;
            const int ic = (int)(w % T.nctr);
            const size_t r = w / T.nctr;
            const double *pp = bp + T.off_p + q + post * ((size_t)ic * nfp + (size_t)T.nctr * nfp * r);
            const double *pm = bm + T.off_m + q + post * ((size_t)ic * nfm + (size_t)T.nctr * nfm * r);
            double v = 0;
            const int a0 = sph ? 0 : m, a1 = sph ? nfi : m + 1;
            for (int a = a0; a < a1; a++) {
                const double coef = sph ? cm[m * nfi + a] : 1.0;
                if (coef == 0.0) continue;
                int ax, ay, az;
                ip_cart_xyz(li, a, ax, ay, az);
                const int n = comp == 0 ? ax : comp == 1 ? ay : az;
                const int up = ip_cart_index(ax + (comp == 0), az + (comp == 2), li + 1);
                double d = sp * pp[post * up];
                if (n > 0 && T.has_m) d += n * sm * pm[post * ip_cart_index(ax - (comp == 0), az - (comp == 2), li - 1)];
                v = fma(coef, d, v);
            }
            out[T.off_o + (size_t)comp * T.comp_stride + q + post * ((size_t)ic * di + m + (size_t)T.nctr * di * r)] = v;
        }
    }
}

CINTOpt *ctx_deriv(CINTOpt *c)
{
    std::lock_guard<std::mutex> lock(c->mtx);
    if (c->deriv) return c->deriv;
    const int nb = c->nbas;
    std::vector<int> xbas(c->bas);
    xbas.resize((size_t)3 * nb * BAS_SLOTS);
    std::vector<double> xenv(c->env);
    for (int i = 0; i < nb; i++) {
        const ShellInfo &s = c->shells[i];
        int *up = xbas.data() + (size_t)(nb + i) * BAS_SLOTS, *dn = xbas.data() + (size_t)(2 * nb + i) * BAS_SLOTS;
        memcpy(up, c->bas.data() + (size_t)i * BAS_SLOTS, sizeof(int) * BAS_SLOTS);
        memcpy(dn, c->bas.data() + (size_t)i * BAS_SLOTS, sizeof(int) * BAS_SLOTS);
        up[ANG_OF] = std::min(s.l + 1, B200_LMAX);          // l = LMAX: placeholder, rejected when it is differentiated
        dn[ANG_OF] = std::max(s.l - 1, 0);                  // l = 0: placeholder, never used
        up[PTR_COEFF] = (int)xenv.size();
        for (int k = 0; k < s.nctr; k++)
            for (int p = 0; p < s.nprim; p++) xenv.push_back(-2.0 * s.exps[p] * s.coef[k * s.nprim + p]);
    }
    CINTOpt *d = NULL;
    if (cintb200_create(&d, c->atm.data(), c->natm, xbas.data(), 3 * nb, xenv.data(), c->device)) return NULL;
    c->deriv = d;
    return d;
}

// dpos: position of the differentiated shell inside the tuple (0 = i: the ip1 integrals; ncenter - 1 = k: int3c2e_ip2, int2c2e_ip2)
static long run_batch_ip(CINTOpt *c, int ncenter, int dpos, int kind, const int *shls, size_t n, const size_t *out_off,
                         double *out, int on_device, int *nonzero)
{
    if (!c || c->magic != B200_CTX_MAGIC) return b200_fail(CINTB200_EINVAL, "invalid context");
    if (n == 0) return 0;
    if (!shls || !out) return b200_fail(CINTB200_EINVAL, "NULL shls/out");
    const int cart = (kind == CINTB200_CART), nb = c->nbas;
    for (size_t t = 0; t < n; t++)
        for (int m = 0; m < ncenter; m++) {
            const int sh = shls[t * ncenter + m];
            if (sh < 0 || sh >= nb) return b200_fail(CINTB200_EINVAL, "tuple %zu: shell id %d out of range", t, sh);
            if (m == dpos && c->shells[sh].l + 1 > B200_LMAX)
                return b200_fail(CINTB200_ENOSUP, "derivative of a shell with l = %d needs l + 1 > %d", c->shells[sh].l, B200_LMAX);
        }
    CINTOpt *d = ctx_deriv(c);
    if (!d) return CINTB200_ENODEV;
    std::vector<int> shp(n * ncenter), shm;
    std::vector<size_t> offp(n), offm;
    std::vector<IpTask> it(n);
    std::vector<size_t> mslot(n, (size_t)-1);
    size_t totp = 0, totm = 0, toto = 0;
    for (size_t t = 0; t < n; t++) {
        const int *s = shls + t * ncenter;
        const ShellInfo &si = c->shells[s[dpos]];
        size_t rest = 1, post = 1;
        for (int m = 0; m < dpos; m++) post *= shell_dim(c->shells[s[m]], cart);
        for (int m = dpos + 1; m < ncenter; m++) rest *= shell_dim(c->shells[s[m]], cart);
        const size_t di = (size_t)(cart ? B200_NCART(si.l) : 2 * si.l + 1) * si.nctr;
        IpTask &T = it[t];
        T.li = si.l; T.nctr = si.nctr; T.rest = (int)rest; T.has_m = si.l > 0; T.post = (int)post; T.pad = 0;
        T.comp_stride = post * di * rest;
        rest *= post;                                   // block sizes below: all other dimensions
        T.off_o = out_off ? out_off[t] : toto;
        toto = std::max(toto, T.off_o + 3 * T.comp_stride);
        for (int m = 0; m < ncenter; m++) shp[t * ncenter + m] = s[m];
        shp[t * ncenter + dpos] = nb + s[dpos];
        offp[t] = T.off_p = totp;
        totp += (size_t)B200_NCART(si.l + 1) * si.nctr * rest;
        T.off_m = 0;
        if (si.l > 0) {
            mslot[t] = offm.size();
            for (int m = 0; m < ncenter; m++) shm.push_back(m == dpos ? 2 * nb + s[m] : s[m]);
            offm.push_back(totm);
            T.off_m = totm;
            totm += (size_t)B200_NCART(si.l - 1) * si.nctr * rest;
        }
    }
    CUDA_OK(cudaSetDevice(c->device));
    double *d_p = nullptr, *d_m = nullptr, *d_o = nullptr;
    IpTask *d_it = nullptr;
    int *d_c2soff = nullptr;
    auto cleanup = [&]() { b200_dfree(d_p); b200_dfree(d_m); b200_dfree(d_it); b200_dfree(d_c2soff); if (!on_device) b200_dfree(d_o); };
    if (b200_dmalloc(&d_p, sizeof(double) * std::max<size_t>(1, totp)) != cudaSuccess || b200_dmalloc(&d_m, sizeof(double) * std::max<size_t>(1, totm)) != cudaSuccess ||
        b200_dmalloc(&d_it, sizeof(IpTask) * n) != cudaSuccess || b200_dmalloc(&d_c2soff, sizeof(C2S_OFF)) != cudaSuccess) {
        cleanup();
        return b200_fail(CINTB200_ENOMEM, "derivative scratch allocation failed");
    }
    d_o = out;
    if (!on_device && b200_dmalloc(&d_o, sizeof(double) * toto) != cudaSuccess) { d_o = nullptr; cleanup(); return b200_fail(CINTB200_ENOMEM, "derivative output allocation failed"); }
    std::vector<int> nzp(n, 0), nzm(offm.size(), 0);
    // Cartesian kind: every index of the helper blocks is Cartesian anyway, so they are ordinary int2e_cart batches and run on
    // the specialised tile kernels (list mode); spherical kind: the differentiated index alone stays Cartesian (generic kernel)
    const int cpos = cart ? -1 : dpos;
    long rc = run_batch(d, ncenter, kind, shp.data(), n, offp.data(), d_p, 1, nzp.data(), cpos);
    if (rc >= 0 && !offm.empty()) rc = run_batch(d, ncenter, kind, shm.data(), offm.size(), offm.data(), d_m, 1, nzm.data(), cpos);
    if (rc < 0) { cleanup(); return rc; }
    {
        std::lock_guard<std::mutex> lock(c->mtx);
        cudaMemcpyAsync(d_it, it.data(), sizeof(IpTask) * n, cudaMemcpyHostToDevice, c->stream);
        cudaMemcpyAsync(d_c2soff, C2S_OFF, sizeof(C2S_OFF), cudaMemcpyHostToDevice, c->stream);
        const unsigned grid = (unsigned)std::min<size_t>(n, 148 * 16);
        ip1_assemble_kernel<<<grid, 128, 0, c->stream>>>(d_it, n, d_p, d_m, d_o, c->d_c2s, d_c2soff, cart);
        c->launches++;
        if (!on_device) cudaMemcpyAsync(out, d_o, sizeof(double) * toto, cudaMemcpyDeviceToHost, c->stream);
        cudaError_t e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { cleanup(); return b200_fail(CINTB200_ENODEV, "derivative assembly failed: %s", cudaGetErrorString(e)); }
    }
    if (nonzero) for (size_t t = 0; t < n; t++) nonzero[t] = nzp[t] | (mslot[t] != (size_t)-1 ? nzm[mslot[t]] : 0);
    cleanup();
    return (long)n;
}

extern "C" long cintb200_int2e_ip1_batch(cintb200_ctx *c, int kind, const int *shls, size_t n, const size_t *out_off,
                                         double *out, int on_device, int *nonzero)
{ return run_batch_ip(c, 4, 0, kind, shls, n, out_off, out, on_device, nonzero); }

extern "C" long cintb200_int3c2e_ip1_batch(cintb200_ctx *c, int kind, const int *shls, size_t n, const size_t *out_off,
                                           double *out, int on_device, int *nonzero)
{ return run_batch_ip(c, 3, 0, kind, shls, n, out_off, out, on_device, nonzero); }

// ( i j | nabla k ), ( nabla i | k ), ( i | nabla k ): src/autocode/int3c2e.c:99-168, :330-383, :408-461
extern "C" long cintb200_int3c2e_ip2_batch(cintb200_ctx *c, int kind, const int *shls, size_t n, const size_t *out_off,
                                           double *out, int on_device, int *nonzero)
{ return run_batch_ip(c, 3, 2, kind, shls, n, out_off, out, on_device, nonzero); }
extern "C" long cintb200_int2c2e_ip1_batch(cintb200_ctx *c, int kind, const int *shls, size_t n, const size_t *out_off,
                                           double *out, int on_device, int *nonzero)
{ return run_batch_ip(c, 2, 0, kind, shls, n, out_off, out, on_device, nonzero); }
extern "C" long cintb200_int2c2e_ip2_batch(cintb200_ctx *c, int kind, const int *shls, size_t n, const size_t *out_off,
                                           double *out, int on_device, int *nonzero)
{ return run_batch_ip(c, 2, 1, kind, shls, n, out_off, out, on_device, nonzero); }

// ------------------------------------------------------------------ Schwarz bounds (device)
// q[p] = sqrt(max |(ij|ij)|) over the block of shell pair p: |(ij|kl)| <= q[ij] q[kl].  The reference has no
// shell-quartet screening (its callers do it); the whole-job driver uses these bounds to skip work items
// whose 32 quartets are all below the threshold (their blocks are zero-filled, like the reference's empty blocks).
__global__ void block_maxabs_kernel(const double *__restrict__ v, const size_t *__restrict__ off, const size_t *__restrict__ len,
                                    double *__restrict__ q, size_t n)
{
    const size_t p = blockIdx.x;
    if (p >= n) return;
    double m = 0;
    for (size_t i = threadIdx.x; i < len[p]; i += blockDim.x) m = fmax(m, fabs(v[off[p] + i]));
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ double sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) m = fmax(m, sm[w]);
        q[p] = sqrt(m);
    }
}

int ctx_compute_schwarz(CINTOpt *c)
{
    if (!c->schwarz.empty()) return 0;
    const size_t np = (size_t)c->nbas * (c->nbas + 1) / 2;
    std::vector<int> shls(4 * np);
    std::vector<size_t> off(np), len(np);
    size_t total = 0;
    for (int i = 0, p = 0; i < c->nbas; i++)
        for (int j = 0; j <= i; j++, p++) {
            shls[4 * p] = i; shls[4 * p + 1] = j; shls[4 * p + 2] = i; shls[4 * p + 3] = j;
            const size_t d = (size_t)(2 * c->shells[i].l + 1) * c->shells[i].nctr * (2 * c->shells[j].l + 1) * c->shells[j].nctr;
            off[p] = total; len[p] = d * d; total += d * d;
        }
    double *d_v = nullptr, *d_q = nullptr;
    size_t *d_off = nullptr, *d_len = nullptr;
    CUDA_OK(cudaSetDevice(c->device));
    CUDA_OK(b200_dmalloc(&d_v, sizeof(double) * total));
    CUDA_OK(b200_dmalloc(&d_q, sizeof(double) * np));
    CUDA_OK(b200_dmalloc(&d_off, sizeof(size_t) * np));
    CUDA_OK(b200_dmalloc(&d_len, sizeof(size_t) * np));
    long rc = run_batch(c, 4, CINTB200_SPH, shls.data(), np, off.data(), d_v, 1, nullptr);
    if (rc >= 0) {
        cudaMemcpy(d_off, off.data(), sizeof(size_t) * np, cudaMemcpyHostToDevice);
        cudaMemcpy(d_len, len.data(), sizeof(size_t) * np, cudaMemcpyHostToDevice);
        block_maxabs_kernel<<<(unsigned)np, 128>>>(d_v, d_off, d_len, d_q, np);
        c->schwarz.resize(np);
        if (cudaMemcpy(c->schwarz.data(), d_q, sizeof(double) * np, cudaMemcpyDeviceToHost) != cudaSuccess) { c->schwarz.clear(); rc = -1; }
    }
    b200_dfree(d_v); b200_dfree(d_q); b200_dfree(d_off); b200_dfree(d_len);
    return rc < 0 ? b200_fail(CINTB200_ENODEV, "Schwarz bound evaluation failed") : 0;
}

extern "C" int cintb200_set_schwarz_threshold(cintb200_ctx *c, double thr)
{
    if (!c || c->magic != B200_CTX_MAGIC) return b200_fail(CINTB200_EINVAL, "invalid context");
    c->schwarz_thr = thr;
    return 0;
}

extern "C" int cintb200_schwarz_bounds(cintb200_ctx *c, double *q)
{
    if (!c || c->magic != B200_CTX_MAGIC) return b200_fail(CINTB200_EINVAL, "invalid context");
    if (ctx_compute_schwarz(c)) return CINTB200_ENODEV;
    if (q) memcpy(q, c->schwarz.data(), sizeof(double) * c->schwarz.size());
    return (int)c->schwarz.size();
}

// ------------------------------------------------------------------ libcint drop-in calls
// Contexts for calls that pass opt == NULL (or an opt built from different arrays) are cached by
// content hash so that loops such as testsuite/test_cint.py:235-256 do not rebuild tables per call.
static std::mutex g_cache_mtx;
static std::vector<CINTOpt *> g_cache;

// RAII holder of a context used by one drop-in call: cached contexts are reference-counted, so that the eviction of the
// least-recently-used entry can never destroy a context another thread is still inside (concurrent callers with
// opt == NULL on different molecules, examples/time_c60.c:196-219 style OpenMP loops).
struct CtxRef {
    CINTOpt *c = nullptr;
    bool counted = false;
    ~CtxRef() { if (c && counted) c->users.fetch_sub(1); }
};

static void context_for(CtxRef &ref, CINTOpt *opt, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env)
{
    const uint64_t h = basis_hash(atm, natm, bas, nbas, env);
    if (opt && opt->magic == B200_CTX_MAGIC && opt->hash == h) { ref.c = opt; return; }      // the caller's own optimizer: caller keeps it alive
    std::lock_guard<std::mutex> lock(g_cache_mtx);
    for (size_t i = 0; i < g_cache.size(); i++)
        if (g_cache[i]->hash == h) {
            CINTOpt *c = g_cache[i];
            g_cache.erase(g_cache.begin() + i);
            g_cache.push_back(c);
            c->users.fetch_add(1);
            ref.c = c; ref.counted = true;
            return;
        }
    CINTOpt *c = NULL;
    if (cintb200_create(&c, atm, natm, bas, nbas, env, -1)) return;
    // evict the least-recently-used IDLE context once more than 8 are cached (busy ones stay: the cache may grow for a while)
    for (size_t i = 0; g_cache.size() >= 8 && i < g_cache.size();) {
        if (g_cache[i]->users.load() == 0) { cintb200_destroy(g_cache[i]); g_cache.erase(g_cache.begin() + i); }
        else i++;
    }
    c->users.fetch_add(1);
    g_cache.push_back(c);
    ref.c = c; ref.counted = true;
}

// failed drop-in call: the reference cannot fail; hand back a zero block (what it does for screened-out blocks) and 0,
// the reason stays in cintb200_last_error() / on stderr
static CACHE_SIZE_T drop_in_failed(double *out, const FINT *dims, const FINT *shls, const FINT *bas, int ncenter, int cart, int ncomp)
{
    if (!out) return 0;
    size_t d[4] = {1, 1, 1, 1};
    for (int m = 0; m < ncenter; m++) d[m] = cart ? CINTcgto_cart(shls[m], bas) : CINTcgto_spheric(shls[m], bas);
    if (!dims) { memset(out, 0, sizeof(double) * ncomp * d[0] * d[1] * d[2] * d[3]); return 0; }
    const size_t ni = dims[0], nj = dims[1], nk = (ncenter > 2) ? dims[2] : 1, nl = (ncenter > 3) ? dims[3] : 1;
    for (int comp = 0; comp < ncomp; comp++)
        for (size_t l = 0; l < d[3]; l++)
            for (size_t k = 0; k < d[2]; k++)
                for (size_t j = 0; j < d[1]; j++)
                    memset(out + comp * ni * nj * nk * nl + ni * (j + nj * (k + nk * l)), 0, sizeof(double) * d[0]);
    return 0;
}

// cart_pos >= 0: that index of a spherical call stays Cartesian (int3c2e_sph_ssc: the auxiliary index, src/cint3c2e.c:729)
static CACHE_SIZE_T drop_in(int ncenter, int kind, double *out, FINT *dims, FINT *shls, FINT *atm, FINT natm,
                            FINT *bas, FINT nbas, double *env, CINTOpt *opt, int cart_pos = -1)
{
    if (out == NULL) {
        // reference: required scratch length in doubles (src/cint2e.c:801-816).  The GPU path needs no
        // caller scratch; report the block size so callers that size a buffer from it stay valid.
        size_t n = 1;
        for (int m = 0; m < ncenter; m++) n *= CINTcgto_cart(shls[m], bas);
        return (CACHE_SIZE_T)n;
    }
    const int cart = (kind == CINTB200_CART);
    for (int m = 0; m < ncenter; m++)
        if (shls[m] < 0 || shls[m] >= nbas) { b200_fail(CINTB200_EINVAL, "shell id %d out of range", shls[m]); return 0; }
    CtxRef ref;
    context_for(ref, opt, atm, natm, bas, nbas, env);
    CINTOpt *c = ref.c;
    if (!c) return drop_in_failed(out, dims, shls, bas, ncenter, cart, 1);
    size_t d[4] = {1, 1, 1, 1};
    for (int m = 0; m < ncenter; m++) d[m] = shell_dim(c->shells[shls[m]], cart || m == cart_pos);
    const size_t len = d[0] * d[1] * d[2] * d[3];
    int nz = 0;
    if (!dims) {
        long rc = run_batch(c, ncenter, kind, shls, 1, NULL, out, 0, &nz, cart_pos);
        return rc < 0 ? drop_in_failed(out, dims, shls, bas, ncenter, cart, 1) : nz;
    }
    std::vector<double> tmp(len);
    long rc = run_batch(c, ncenter, kind, shls, 1, NULL, tmp.data(), 0, &nz, cart_pos);
    if (rc < 0) return drop_in_failed(out, dims, shls, bas, ncenter, cart, 1);
    // embed into the caller's larger tensor: leading dimensions dims[] (src/cint2e.c:853-856)
    const size_t ni = dims[0], nj = dims[1], nk = (ncenter > 2) ? dims[2] : 1;
    for (size_t l = 0; l < d[3]; l++)
        for (size_t k = 0; k < d[2]; k++)
            for (size_t j = 0; j < d[1]; j++)
                memcpy(out + ni * (j + nj * (k + nk * l)), tmp.data() + d[0] * (j + d[1] * (k + d[2] * l)), sizeof(double) * d[0]);
    return nz;
}

// ( nabla i j | k l ): three blocks, out[comp][l][k][j][i]; with dims the stride between components is the product of dims
static CACHE_SIZE_T drop_in_ip1(int ncenter, int kind, double *out, FINT *dims, FINT *shls, FINT *atm, FINT natm,
                                FINT *bas, FINT nbas, double *env, CINTOpt *opt, int dpos = 0)
{
    if (out == NULL) {
        size_t n = 3;
        for (int m = 0; m < ncenter; m++) n *= CINTcgto_cart(shls[m], bas);
        return (CACHE_SIZE_T)n;
    }
    const int cart = (kind == CINTB200_CART);
    for (int m = 0; m < ncenter; m++)
        if (shls[m] < 0 || shls[m] >= nbas) { b200_fail(CINTB200_EINVAL, "shell id %d out of range", shls[m]); return 0; }
    CtxRef ref;
    context_for(ref, opt, atm, natm, bas, nbas, env);
    CINTOpt *c = ref.c;
    if (!c) return drop_in_failed(out, dims, shls, bas, ncenter, cart, 3);
    size_t d[4] = {1, 1, 1, 1};
    for (int m = 0; m < ncenter; m++) d[m] = shell_dim(c->shells[shls[m]], cart);
    const size_t len = d[0] * d[1] * d[2] * d[3];
    int nz = 0;
    if (!dims) {
        long rc = run_batch_ip(c, ncenter, dpos, kind, shls, 1, NULL, out, 0, &nz);
        return rc < 0 ? drop_in_failed(out, dims, shls, bas, ncenter, cart, 3) : nz;
    }
    std::vector<double> tmp(3 * len);
    long rc = run_batch_ip(c, ncenter, dpos, kind, shls, 1, NULL, tmp.data(), 0, &nz);
    if (rc < 0) return drop_in_failed(out, dims, shls, bas, ncenter, cart, 3);
    const size_t ni = dims[0], nj = dims[1], nk = (ncenter > 2) ? dims[2] : 1, nl = (ncenter > 3) ? dims[3] : 1;
    for (size_t comp = 0; comp < 3; comp++)
        for (size_t l = 0; l < d[3]; l++)
            for (size_t k = 0; k < d[2]; k++)
                for (size_t j = 0; j < d[1]; j++)
                    memcpy(out + comp * ni * nj * nk * nl + ni * (j + nj * (k + nk * l)),
                           tmp.data() + comp * len + d[0] * (j + d[1] * (k + d[2] * l)), sizeof(double) * d[0]);
    return nz;
}

extern "C" {
CACHE_SIZE_T int2e_ip1_sph(double *out, FINT *dims, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt, double *cache)
{ (void)cache; return drop_in_ip1(4, CINTB200_SPH, out, dims, shls, atm, natm, bas, nbas, env, opt); }
CACHE_SIZE_T int2e_ip1_cart(double *out, FINT *dims, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt, double *cache)
{ (void)cache; return drop_in_ip1(4, CINTB200_CART, out, dims, shls, atm, natm, bas, nbas, env, opt); }
CACHE_SIZE_T int3c2e_ip1_sph(double *out, FINT *dims, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt, double *cache)
{ (void)cache; return drop_in_ip1(3, CINTB200_SPH, out, dims, shls, atm, natm, bas, nbas, env, opt); }
CACHE_SIZE_T int3c2e_ip1_cart(double *out, FINT *dims, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt, double *cache)
{ (void)cache; return drop_in_ip1(3, CINTB200_CART, out, dims, shls, atm, natm, bas, nbas, env, opt); }
CACHE_SIZE_T int3c2e_ip2_sph(double *out, FINT *dims, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt, double *cache)
{ (void)cache; return drop_in_ip1(3, CINTB200_SPH, out, dims, shls, atm, natm, bas, nbas, env, opt, 2); }
CACHE_SIZE_T int3c2e_ip2_cart(double *out, FINT *dims, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt, double *cache)
{ (void)cache; return drop_in_ip1(3, CINTB200_CART, out, dims, shls, atm, natm, bas, nbas, env, opt, 2); }
CACHE_SIZE_T int2c2e_ip1_sph(double *out, FINT *dims, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt, double *cache)
{ (void)cache; return drop_in_ip1(2, CINTB200_SPH, out, dims, shls, atm, natm, bas, nbas, env, opt, 0); }
CACHE_SIZE_T int2c2e_ip2_sph(double *out, FINT *dims, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt, double *cache)
{ (void)cache; return drop_in_ip1(2, CINTB200_SPH, out, dims, shls, atm, natm, bas, nbas, env, opt, 1); }
void int3c2e_ip2_optimizer(CINTOpt **opt, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env)
{ *opt = NULL; cintb200_create(opt, atm, natm, bas, nbas, env, -1); }
void int2c2e_ip1_optimizer(CINTOpt **opt, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env)
{ *opt = NULL; cintb200_create(opt, atm, natm, bas, nbas, env, -1); }
void int2c2e_ip2_optimizer(CINTOpt **opt, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env)
{ *opt = NULL; cintb200_create(opt, atm, natm, bas, nbas, env, -1); }
FINT cint3c2e_ip2_sph(double *out, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt)
{ return int3c2e_ip2_sph(out, NULL, shls, atm, natm, bas, nbas, env, opt, NULL); }
FINT cint2c2e_ip1_sph(double *out, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt)
{ return int2c2e_ip1_sph(out, NULL, shls, atm, natm, bas, nbas, env, opt, NULL); }
FINT cint2c2e_ip2_sph(double *out, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt)
{ return int2c2e_ip2_sph(out, NULL, shls, atm, natm, bas, nbas, env, opt, NULL); }
void int2e_ip1_optimizer(CINTOpt **opt, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env)
{ *opt = NULL; cintb200_create(opt, atm, natm, bas, nbas, env, -1); }
void int3c2e_ip1_optimizer(CINTOpt **opt, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env)
{ *opt = NULL; cintb200_create(opt, atm, natm, bas, nbas, env, -1); }
FINT cint2e_ip1_sph(double *out, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt)
{ return int2e_ip1_sph(out, NULL, shls, atm, natm, bas, nbas, env, opt, NULL); }
FINT cint3c2e_ip1_sph(double *out, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt)
{ return int3c2e_ip1_sph(out, NULL, shls, atm, natm, bas, nbas, env, opt, NULL); }
void cint2e_ip1_sph_optimizer(CINTOpt **opt, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env) { int2e_ip1_optimizer(opt, atm, natm, bas, nbas, env); }
void cint3c2e_ip1_sph_optimizer(CINTOpt **opt, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env) { int3c2e_ip1_optimizer(opt, atm, natm, bas, nbas, env); }

CACHE_SIZE_T int2e_sph(double *out, FINT *dims, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt, double *cache)
{ (void)cache; return drop_in(4, CINTB200_SPH, out, dims, shls, atm, natm, bas, nbas, env, opt); }
CACHE_SIZE_T int2e_cart(double *out, FINT *dims, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt, double *cache)
{ (void)cache; return drop_in(4, CINTB200_CART, out, dims, shls, atm, natm, bas, nbas, env, opt); }
CACHE_SIZE_T int3c2e_sph(double *out, FINT *dims, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt, double *cache)
{ (void)cache; return drop_in(3, CINTB200_SPH, out, dims, shls, atm, natm, bas, nbas, env, opt); }
CACHE_SIZE_T int3c2e_cart(double *out, FINT *dims, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt, double *cache)
{ (void)cache; return drop_in(3, CINTB200_CART, out, dims, shls, atm, natm, bas, nbas, env, opt); }
// spherical i, j with a Cartesian auxiliary index k (c2s_sph_3c2e1_ssc, src/cart2sph.c:5956; entry point src/cint3c2e.c:729)
CACHE_SIZE_T int3c2e_sph_ssc(double *out, FINT *dims, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt, double *cache)
{ (void)cache; return drop_in(3, CINTB200_SPH, out, dims, shls, atm, natm, bas, nbas, env, opt, 2); }
void int3c2e_ssc_optimizer(CINTOpt **opt, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env) { int3c2e_optimizer(opt, atm, natm, bas, nbas, env); }

CACHE_SIZE_T int2c2e_sph(double *out, FINT *dims, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt, double *cache)
{ (void)cache; return drop_in(2, CINTB200_SPH, out, dims, shls, atm, natm, bas, nbas, env, opt); }
CACHE_SIZE_T int2c2e_cart(double *out, FINT *dims, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt, double *cache)
{ (void)cache; return drop_in(2, CINTB200_CART, out, dims, shls, atm, natm, bas, nbas, env, opt); }
void int2c2e_optimizer(CINTOpt **opt, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env)
{ *opt = NULL; cintb200_create(opt, atm, natm, bas, nbas, env, -1); }
FINT cint2c2e_sph(double *out, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt)
{ return int2c2e_sph(out, NULL, shls, atm, natm, bas, nbas, env, opt, NULL); }
FINT cint2c2e_cart(double *out, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt)
{ return int2c2e_cart(out, NULL, shls, atm, natm, bas, nbas, env, opt, NULL); }
void cint2c2e_sph_optimizer(CINTOpt **opt, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env) { int2c2e_optimizer(opt, atm, natm, bas, nbas, env); }
void cint2c2e_cart_optimizer(CINTOpt **opt, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env) { int2c2e_optimizer(opt, atm, natm, bas, nbas, env); }

void int2e_optimizer(CINTOpt **opt, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env)
{ *opt = NULL; cintb200_create(opt, atm, natm, bas, nbas, env, -1); }
void int3c2e_optimizer(CINTOpt **opt, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env)
{ *opt = NULL; cintb200_create(opt, atm, natm, bas, nbas, env, -1); }

FINT cint2e_sph(double *out, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt)
{ return int2e_sph(out, NULL, shls, atm, natm, bas, nbas, env, opt, NULL); }
FINT cint2e_cart(double *out, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt)
{ return int2e_cart(out, NULL, shls, atm, natm, bas, nbas, env, opt, NULL); }
FINT cint3c2e_sph(double *out, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt)
{ return int3c2e_sph(out, NULL, shls, atm, natm, bas, nbas, env, opt, NULL); }
FINT cint3c2e_cart(double *out, FINT *shls, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env, CINTOpt *opt)
{ return int3c2e_cart(out, NULL, shls, atm, natm, bas, nbas, env, opt, NULL); }
void cint2e_sph_optimizer(CINTOpt **opt, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env) { int2e_optimizer(opt, atm, natm, bas, nbas, env); }
void cint2e_cart_optimizer(CINTOpt **opt, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env) { int2e_optimizer(opt, atm, natm, bas, nbas, env); }
void cint3c2e_sph_optimizer(CINTOpt **opt, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env) { int3c2e_optimizer(opt, atm, natm, bas, nbas, env); }
void cint3c2e_cart_optimizer(CINTOpt **opt, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env) { int3c2e_optimizer(opt, atm, natm, bas, nbas, env); }

// src/optimizer.c:22-72: init = "empty optimizer".  An empty optimizer carries no tables, so NULL
// (which every integral entry point accepts) is the faithful equivalent.
void CINTinit_2e_optimizer(CINTOpt **opt, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env)
{ (void)atm; (void)natm; (void)bas; (void)nbas; (void)env; *opt = NULL; }
void CINTinit_optimizer(CINTOpt **opt, FINT *atm, FINT natm, FINT *bas, FINT nbas, double *env)
{ CINTinit_2e_optimizer(opt, atm, natm, bas, nbas, env); }
void CINTdel_2e_optimizer(CINTOpt **opt)
{ if (opt && *opt) { cintb200_destroy(*opt); *opt = NULL; } }
void CINTdel_optimizer(CINTOpt **opt) { CINTdel_2e_optimizer(opt); }
}

const int *engine_c2s_off() { return C2S_OFF; }
extern "C" void cintb200_debug_force_generic(cintb200_ctx *c, int on) { if (c && c->magic == B200_CTX_MAGIC) c->force_generic = on; }

int rys_tab_nint(int nroots) { return (nroots >= 1 && nroots <= RYS_NMAX) ? RYS_TAB_NINT[nroots] : 0; }
int rys_tab_off(int nroots) { return (nroots >= 1 && nroots <= RYS_NMAX) ? RYS_TAB_OFF[nroots] : 0; }
int rys_fast_nint(int nroots) { return (nroots >= 1 && nroots <= RYS_FNMAX) ? RYS_FAST_NINT[nroots] : 0; }
int rys_fast_off(int nroots) { return (nroots >= 1 && nroots <= RYS_FNMAX) ? RYS_FAST_OFF[nroots] : 0; }

// ------------------------------------------------------------------ large device buffers
// Tile buffers and digestion partials come from the device's stream-ordered memory pool with the release threshold raised:
// a destroyed context's buffers are then reused by the next context instead of being unmapped and mapped again (measured:
// cudaFree of the 80 GB tile buffer 0.94 s, cudaMalloc ~0.1 s -- per context, which is per geometry step for a caller).
// cintb200_release_cached_memory() hands the cached memory back to the driver.
static std::mutex g_big_mtx;
static std::map<void *, int> g_big_pooled;            // pointer -> device, for buffers that came from the pool
static std::map<int, cudaStream_t> g_big_stream;

static cudaStream_t big_stream(int dev)
{
    auto it = g_big_stream.find(dev);
    if (it != g_big_stream.end()) return it->second;
    cudaStream_t s = nullptr;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) s = nullptr;
    }
    cudaGetLastError();
    g_big_stream[dev] = s;
    return s;
}

int b200_big_alloc(void **p, size_t bytes)
{
    std::lock_guard<std::mutex> lock(g_big_mtx);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaStream_t s = big_stream(dev);
    *p = nullptr;
    for (int attempt = 0; s && attempt < 2; attempt++) {
        if (cudaMallocAsync(p, bytes ? bytes : 1, s) == cudaSuccess && cudaStreamSynchronize(s) == cudaSuccess) {
            g_big_pooled[*p] = dev;
            return 0;
        }
        cudaGetLastError();
        *p = nullptr;
        cudaMemPool_t pool;                         // out of memory with cached blocks of the wrong sizes: trim and retry once
        if (cudaDeviceGetDefaultMemPool(&pool, dev) != cudaSuccess) break;
        cudaDeviceSynchronize();
        cudaMemPoolTrimTo(pool, 0);
    }
    cudaGetLastError();
    *p = nullptr;
    return cudaMalloc(p, bytes ? bytes : 1) == cudaSuccess ? 0 : -1;
}

void b200_big_free(void *p)
{
    if (!p) return;
    std::lock_guard<std::mutex> lock(g_big_mtx);
    auto it = g_big_pooled.find(p);
    if (it == g_big_pooled.end()) { cudaFree(p); return; }
    cudaStream_t s = big_stream(it->second);
    g_big_pooled.erase(it);
    cudaDeviceSynchronize();                        // every consumer of the buffer has finished (callers free at teardown only)
    cudaFreeAsync(p, s);
    cudaStreamSynchronize(s);
}

extern "C" int cintb200_release_cached_memory(int device)
{
    std::lock_guard<std::mutex> lock(g_big_mtx);
    int dev = device;
    if (dev < 0) cudaGetDevice(&dev);
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) != cudaSuccess) return b200_fail(CINTB200_ENODEV, "no memory pool on device %d", dev);
    cudaDeviceSynchronize();
    return cudaMemPoolTrimTo(pool, 0) == cudaSuccess ? 0 : b200_fail(CINTB200_ENODEV, "cudaMemPoolTrimTo failed");
}

// ------------------------------------------------------------------ launch geometry of the persistent tile kernels
int tile_smem_limit()
{
    static int lim = 0;
    if (!lim) {
        int dev = 0, v = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || v <= 0) v = 48 * 1024;
        lim = v;
    }
    return lim;
}

int tile_grid_blocks(const void *fn, int threads, size_t smem, long long total)
{
    struct Info { size_t smem_set = 0; int occ = 0; };
    static std::mutex mtx;
    static std::map<std::pair<int, const void *>, Info> cache;      // function attributes are per device
    static std::map<int, int> sm_count;
    static int per_sm = -1;                     // 0 = occupancy-sized grid, > 0 = that many blocks per SM
    std::lock_guard<std::mutex> lock(mtx);
    int dev = 0;
    cudaGetDevice(&dev);
    if (per_sm < 0) {
        const char *e = getenv("CINTB200_PBLOCKS");
        per_sm = (e && atoi(e) > 0) ? atoi(e) : 0;       // measured on C60: occupancy-sized 1137 ms, 4/SM 1143, 8/SM 1150, 16/SM 1156, 32/SM 1171
    }
    int &sms = sm_count[dev];
    if (!sms && (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)) sms = 148;
    Info &in = cache[std::make_pair(dev, fn)];
    if (smem > in.smem_set || !in.smem_set) {
        int lim = 0;
        if (cudaDeviceGetAttribute(&lim, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || lim <= 0) lim = 48 * 1024;
        if (smem > (size_t)lim) return -1;
        if (smem > 48 * 1024 && cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
        in.smem_set = smem > 0 ? smem : 1;
        in.occ = 0;
    }
    int blocks_per_sm = per_sm;
    if (per_sm == 0) {
        if (!in.occ && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&in.occ, fn, threads, smem) != cudaSuccess) in.occ = 1;
        blocks_per_sm = in.occ > 0 ? in.occ : 1;
    }
    const long long cap = (long long)blocks_per_sm * sms;
    return (int)(total < cap ? (total > 0 ? total : 1) : cap);
}

// ------------------------------------------------------------------ FP64 roofline denominator
// MEASURED_PEAKS.json carries no FP64 entry, so the bench measures the DFMA peak itself: 16 independent FMA chains per
// thread, the loop unrolled 16x (one loop-control sequence per 256 DFMAs), 2 flops per FMA, 8 warps per scheduler.
// cintb200_fp64_peak_theoretical gives SMs x 64 FP64 lanes x 2 x the SM clock for comparison.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double a, double b)
{
    double v[16];
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = threadIdx.x + k;
#pragma unroll 16
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) v[k] = fma(v[k], a, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) s += v[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" int cintb200_fp64_peak_theoretical(int device, double sm_mhz, double *tflops)
{
    if (device >= 0) CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    int dev;
    CUDA_OK(cudaGetDevice(&dev));
    CUDA_OK(cudaGetDeviceProperties(&prop, dev));
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const double mhz = sm_mhz > 0 ? sm_mhz : khz * 1e-3;
    *tflops = prop.multiProcessorCount * 64.0 * 2.0 * mhz * 1e6 / 1e12;        // 64 FP64 FMA lanes per SM on sm_100
    return 0;
}

extern "C" int cintb200_fp64_peak(int device, double seconds, double *tflops)
{
    if (device >= 0) CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    int dev;
    CUDA_OK(cudaGetDevice(&dev));
    CUDA_OK(cudaGetDeviceProperties(&prop, dev));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
    double *buf;
    CUDA_OK(b200_dmalloc(&buf, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0));
    CUDA_OK(cudaEventCreate(&e1));
    dfma_peak_kernel<<<blocks, threads>>>(buf, iters, 0.999999, 1e-9);      // warm-up
    CUDA_OK(cudaDeviceSynchronize());
    double best = 0, spent = 0;
    while (spent < seconds) {
        CUDA_OK(cudaEventRecord(e0));
        dfma_peak_kernel<<<blocks, threads>>>(buf, iters, 0.999999, 1e-9);
        CUDA_OK(cudaEventRecord(e1));
        CUDA_OK(cudaEventSynchronize(e1));
        float ms;
        CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
        const double tf = 2.0 * 16 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
        spent += ms * 1e-3;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); b200_dfree(buf);
    *tflops = best;
    return 0;
}
