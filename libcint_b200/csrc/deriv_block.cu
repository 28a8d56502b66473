// First derivatives on dense shell-slice blocks:  ( nabla_i  i j | k l )  and  ( nabla_i  i j | k ),  three components,
//     out[i + NI (j + NJ (k + NK (l + NL comp)))]        (int2e_ip1, src/autocode/grad2.c:19-68; int3c2e_ip1, src/autocode/int3c2e.c)
// evaluated by the SPECIALISED tile kernels.  The reference differentiates on its g array,
//     d/dX [ (x-X)^n e^{-a (x-X)^2} ] = 2a (x-X)^{n+1} e^{..} - n (x-X)^{n-1} e^{..}            (CINTnabla1i_2e, src/g2e.c:4550)
// Here the identity is applied one level up, on whole shells (engine.cu:ctx_deriv holds, for every shell, a RAISED copy with
// l+1 and coefficients -2 a c, and a LOWERED copy with l-1):
//   1. P = Cartesian block over (raised i-slice, j, k, l), M = the same over the lowered i-slice   -> two dense-block jobs on the
//      tile kernels (driver.cu:run_block with Cartesian output: register / cooperative kernels, CART instantiations);
//   2. derivative in the Cartesian basis, component by component (gather + 2 FMAs per element);
//   3. cart -> sph on each of the four indices of the dense tensor (block-diagonal transforms along one axis at a time).
// The batch entry points (engine.cu:run_batch_ip) keep serving arbitrary tuple lists; this is the throughput path for
// callers that fill AO blocks.
#include <cstdio>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include "../../include/cint_b200.h"
#include "types.h"
#include "engine.h"
#include "c2s_tables.inc"

#define CU_OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    { rc = b200_fail(CINTB200_ENODEV, "%s failed: %s", #call, cudaGetErrorString(e_)); goto done; } } while (0)

// Dc[a + NA (r + R comp)] = cP[a] P[ipP[comp][a] + NP r] + cM[comp][a] M[ipM[comp][a] + NM r]
__global__ void ip1_block_assemble_kernel(const double *__restrict__ P, const double *__restrict__ M, long long NP, long long NM, long long NA, long long R,
                                          const int *__restrict__ ipP, const int *__restrict__ ipM, const double *__restrict__ cP,
                                          const double *__restrict__ cM, double *__restrict__ D)
{
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= NA * R) return;
    const long long a = n % NA, r = n / NA;
#pragma unroll
    for (int comp = 0; comp < 3; comp++) {
        double v = cP[a] * P[ipP[comp * NA + a] + NP * r];
        const int im = ipM[comp * NA + a];
        if (im >= 0) v = fma(cM[comp * NA + a], M[im + NM * r], v);
        D[a + NA * (r + R * comp)] = v;
    }
}

// block-diagonal transform along one axis: out[p + pre (m + nout q)] = sum_{a < cnt[m]} coef[cof[m] + a] in[p + pre (a0[m] + a + nin q)]
__global__ void axis_transform_kernel(const double *__restrict__ in, double *__restrict__ out, long long pre, long long nin, long long nout, long long post,
                                      const int *__restrict__ a0, const int *__restrict__ cnt, const int *__restrict__ cof, const double *__restrict__ coef)
{
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= pre * nout * post) return;
    const long long p = n % pre, m = (n / pre) % nout, q = n / (pre * nout);
    const double *src = in + p + pre * (a0[m] + nin * q);
    const double *cf = coef + cof[m];
    double s = 0;
    for (int a = 0; a < cnt[m]; a++) s = fma(cf[a], src[pre * a], s);
    out[n] = s;
}

static inline int ncart(int l) { return (l + 1) * (l + 2) / 2; }
static inline int cart_index(int lx, int lz, int l) { const int r = l - lx; return r * (r + 1) / 2 + lz; }
static inline void cart_xyz(int l, int idx, int *lx, int *ly, int *lz)
{
    int n = 0;
    for (int x = l; x >= 0; x--) {
        const int cntx = l - x + 1;
        if (idx < n + cntx) { *lx = x; *lz = idx - n; *ly = l - x - *lz; return; }
        n += cntx;
    }
    *lx = *ly = *lz = 0;
}

template <class T>
static int dev_upload(T **dst, const std::vector<T> &v)
{
    if (b200_dmalloc((void **)dst, sizeof(T) * std::max<size_t>(1, v.size())) != cudaSuccess) return -1;
    if (!v.empty() && cudaMemcpy(*dst, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice) != cudaSuccess) return -1;
    return 0;
}

static int ip1_block(cintb200_ctx *c, int ncenter, int kind, const int *sl, double *out, int on_device, double *stats)
{
    if (!c || c->magic != B200_CTX_MAGIC) return b200_fail(CINTB200_EINVAL, "invalid context");
    if (!sl || !out) return b200_fail(CINTB200_EINVAL, "NULL shls_slice/out");
    const int nb = c->nbas, cart = (kind == CINTB200_CART);
    for (int m = 0; m < ncenter; m++)
        if (sl[2 * m] < 0 || sl[2 * m + 1] > nb || sl[2 * m] >= sl[2 * m + 1])
            return b200_fail(CINTB200_EINVAL, "shell slice %d = [%d, %d) is empty or outside 0..%d", m, sl[2 * m], sl[2 * m + 1], nb);
    for (int i = sl[0]; i < sl[1]; i++)
        if (c->shells[i].l + 1 > B200_LMAX) return b200_fail(CINTB200_ENOSUP, "derivative of a shell with l = %d needs l + 1 > %d", c->shells[i].l, B200_LMAX);
    CINTOpt *d = ctx_deriv(c);
    if (!d) return CINTB200_ENODEV;
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    if (cudaSetDevice(c->device) != cudaSuccess) return b200_fail(CINTB200_ENODEV, "cudaSetDevice failed");
    // ---- dimensions ----
    long long NC[4] = {0, 1, 1, 1}, NS[4] = {0, 1, 1, 1};          // Cartesian / spherical AO counts of the four slices
    for (int m = 0; m < ncenter; m++) {
        NC[m] = NS[m] = 0;
        for (int s = sl[2 * m]; s < sl[2 * m + 1]; s++) { NC[m] += (long long)ncart(c->shells[s].l) * c->shells[s].nctr; NS[m] += (long long)(2 * c->shells[s].l + 1) * c->shells[s].nctr; }
    }
    long long NP = 0, NM = 0;
    for (int s = sl[0]; s < sl[1]; s++) {
        NP += (long long)ncart(c->shells[s].l + 1) * c->shells[s].nctr;
        NM += (long long)ncart(std::max(c->shells[s].l - 1, 0)) * c->shells[s].nctr;
    }
    const long long R = NC[1] * NC[2] * NC[3];
    // ---- index / coefficient tables of the Cartesian derivative along the i axis ----
    const double fsp[2] = {0.282094791773878143, 0.488602511902919921};
    std::vector<int> ipP(3 * NC[0]), ipM(3 * NC[0]);
    std::vector<double> cP(NC[0]), cM(3 * NC[0]);
    {
        long long a0 = 0, p0 = 0, m0 = 0;
        for (int s = sl[0]; s < sl[1]; s++) {
            const int l = c->shells[s].l, nf = ncart(l), nfp = ncart(l + 1), nfm = ncart(std::max(l - 1, 0));
            // the engine scales s and p functions by fac_sp(l) instead of transforming them: undo the factor of the raised /
            // lowered shell and apply the one of the target shell (src/g1e.c:565-572)
            const double fi = l < 2 ? fsp[l] : 1.0;
            const double sp = fi / (l + 1 < 2 ? fsp[l + 1] : 1.0), sm = l > 0 ? fi / (l - 1 < 2 ? fsp[l - 1] : 1.0) : 0.0;
            for (int ic = 0; ic < c->shells[s].nctr; ic++)
                for (int f = 0; f < nf; f++) {
                    int ax, ay, az;
                    cart_xyz(l, f, &ax, &ay, &az);
                    const long long a = a0 + (long long)ic * nf + f;
                    cP[a] = sp;
                    for (int comp = 0; comp < 3; comp++) {
                        const int n = comp == 0 ? ax : comp == 1 ? ay : az;
                        ipP[comp * NC[0] + a] = (int)(p0 + (long long)ic * nfp + cart_index(ax + (comp == 0), az + (comp == 2), l + 1));
                        ipM[comp * NC[0] + a] = n > 0 ? (int)(m0 + (long long)ic * nfm + cart_index(ax - (comp == 0), az - (comp == 2), l - 1)) : -1;
                        cM[comp * NC[0] + a] = n * sm;
                    }
                }
            a0 += (long long)nf * c->shells[s].nctr; p0 += (long long)nfp * c->shells[s].nctr; m0 += (long long)nfm * c->shells[s].nctr;
        }
    }
    int rc = 0;
    double *dP = nullptr, *dM = nullptr, *dA = nullptr, *dB = nullptr, *d_coef = nullptr, *d_cP = nullptr, *d_cM = nullptr;
    int *d_ipP = nullptr, *d_ipM = nullptr, *d_tab[4][3] = {{nullptr}};
    const size_t szP = (size_t)(NP * R), szM = (size_t)(NM * R), szD = (size_t)(3 * NC[0] * R);
    double st_p[16] = {0}, st_m[16] = {0};
    cudaStream_t st = c->stream;
    double tph = b200_now();
    {
        std::lock_guard<std::mutex> lock(c->mtx);
        // the four work tensors live in the context and only grow (a gradient loop calls this once per ket shell: allocating
        // them per call cost 9 ms of 35 ms per call on C2H6 cc-pVQZ)
        const size_t want[4] = {szP, szM, szD, cart ? 0 : szD};
        for (int k = 0; k < 4; k++) {
            const size_t bytes = sizeof(double) * std::max<size_t>(1, want[k]);
            if (c->cap_ipwork[k] < bytes) {
                b200_dfree(c->d_ipwork[k]); c->d_ipwork[k] = nullptr; c->cap_ipwork[k] = 0;
                if (b200_big_alloc(&c->d_ipwork[k], bytes)) {
                    rc = b200_fail(CINTB200_ENOMEM, "derivative block: %zu bytes of work tensors", sizeof(double) * (szP + szM + 2 * szD));
                    goto done;
                }
                c->cap_ipwork[k] = bytes;
            }
        }
        dP = (double *)c->d_ipwork[0]; dM = (double *)c->d_ipwork[1]; dA = (double *)c->d_ipwork[2]; dB = (double *)c->d_ipwork[3];
        if (dev_upload(&d_ipP, ipP) || dev_upload(&d_ipM, ipM) || dev_upload(&d_cP, cP) || dev_upload(&d_cM, cM)) { rc = b200_fail(CINTB200_ENOMEM, "derivative block tables"); goto done; }
    }
    b200_phase("ip1 block: work tensors + derivative tables", tph);
    tph = b200_now();
    {
        // 1. the two Cartesian helper blocks on the tile kernels
        int slp[8], slm[8];
        for (int m = 0; m < 2 * ncenter; m++) slp[m] = slm[m] = sl[m];
        slp[0] = nb + sl[0]; slp[1] = nb + sl[1]; slm[0] = 2 * nb + sl[0]; slm[1] = 2 * nb + sl[1];
        rc = run_block(d, ncenter, slp, dP, 1, st_p, 1);
        if (!rc) rc = run_block(d, ncenter, slm, dM, 1, st_m, 1);
        if (rc) goto done;
    }
    b200_phase("ip1 block: helper blocks (plan + kernels)", tph);
    tph = b200_now();
    {
        std::lock_guard<std::mutex> lock(c->mtx);
        // 2. derivative in the Cartesian basis
        const long long nel = NC[0] * R;
        // the last pass of the chain writes straight into a device-resident `out`
        double *cur = (cart && on_device) ? out : dA, *nxt = dB;
        ip1_block_assemble_kernel<<<(unsigned)((nel + 255) / 256), 256, 0, st>>>(dP, dM, NP, NM, NC[0], R, d_ipP, d_ipM, d_cP, d_cM, cur);
        long long dims[5] = {NC[0], NC[1], NC[2], NC[3], 3};
        if (!cart) {
            // 3. cart -> sph along each index (s and p shells: identity; d and higher: the matrices of c2s_tables.inc)
            if (b200_dmalloc((void **)&d_coef, sizeof(double) * (sizeof(C2S_COEF) / sizeof(double) + 1)) != cudaSuccess) { rc = b200_fail(CINTB200_ENOMEM, "c2s table"); goto done; }
            std::vector<double> coef(C2S_COEF, C2S_COEF + sizeof(C2S_COEF) / sizeof(double));
            const int one_off = (int)coef.size();
            coef.push_back(1.0);
            CU_OK(cudaMemcpyAsync(d_coef, coef.data(), sizeof(double) * coef.size(), cudaMemcpyHostToDevice, st));
            CU_OK(cudaStreamSynchronize(st));
            for (int m = 0; m < ncenter; m++) {
                std::vector<int> a0v, cntv, cofv;
                int a0 = 0;
                for (int s = sl[2 * m]; s < sl[2 * m + 1]; s++) {
                    const int l = c->shells[s].l, nf = ncart(l);
                    for (int ic = 0; ic < c->shells[s].nctr; ic++, a0 += nf)
                        for (int ms = 0; ms < (l < 2 ? nf : 2 * l + 1); ms++) {
                            if (l < 2) { a0v.push_back(a0 + ms); cntv.push_back(1); cofv.push_back(one_off); }
                            else { a0v.push_back(a0); cntv.push_back(nf); cofv.push_back(C2S_OFF[l] + ms * nf); }
                        }
                }
                if (dev_upload(&d_tab[m][0], a0v) || dev_upload(&d_tab[m][1], cntv) || dev_upload(&d_tab[m][2], cofv)) { rc = b200_fail(CINTB200_ENOMEM, "c2s axis tables"); goto done; }
                long long pre = 1, post = 1;
                for (int k = 0; k < m; k++) pre *= dims[k];
                for (int k = m + 1; k < 5; k++) post *= dims[k];
                const long long nout = NS[m], tot = pre * nout * post;
                double *dst = (on_device && m == ncenter - 1) ? out : nxt;
                axis_transform_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(cur, dst, pre, dims[m], nout, post, d_tab[m][0], d_tab[m][1], d_tab[m][2], d_coef);
                dims[m] = nout;
                if (dst == out) cur = out; else std::swap(cur, nxt);
            }
        }
        const size_t nout_total = (size_t)(dims[0] * dims[1] * dims[2] * dims[3] * 3);
        if (cur != out) CU_OK(cudaMemcpyAsync(out, cur, sizeof(double) * nout_total, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
        CU_OK(cudaStreamSynchronize(st));
        {
            cudaError_t le = cudaGetLastError();
            if (le != cudaSuccess) { rc = b200_fail(CINTB200_ENODEV, "derivative block kernels failed: %s", cudaGetErrorString(le)); goto done; }
        }
        if (stats) {
            for (int k = 0; k < 12; k++) stats[k] = st_p[k] + st_m[k];
            stats[1] = (double)nout_total;
            stats[9] = 1; stats[10] = sizeof(double) * (double)(szP + szM + 2 * szD);
        }
        c->launches += 1 + (cart ? 0 : ncenter);
    }
done:
    cudaStreamSynchronize(st);
    b200_phase("ip1 block: assemble + c2s + copy", tph);
    tph = b200_now();
    b200_dfree(d_coef); b200_dfree(d_cP); b200_dfree(d_cM); b200_dfree(d_ipP); b200_dfree(d_ipM);
    for (int m = 0; m < 4; m++) for (int k = 0; k < 3; k++) b200_dfree(d_tab[m][k]);
    b200_phase("ip1 block: release", tph);
    if (prev_dev >= 0) cudaSetDevice(prev_dev);
    return rc;
}

extern "C" int cintb200_int2e_ip1_block(cintb200_ctx *c, int kind, const int *shls_slice, double *out, int on_device, double *stats)
{ return ip1_block(c, 4, kind, shls_slice, out, on_device, stats); }
extern "C" int cintb200_int3c2e_ip1_block(cintb200_ctx *c, int kind, const int *shls_slice, double *out, int on_device, double *stats)
{ return ip1_block(c, 3, kind, shls_slice, out, on_device, stats); }
