// Whole-job / tile driver: evaluates every unique shell quartet of the reference benchmark loop
// (examples/time_c60.c:200-219: i>=j, k>=l, k<=i) class by class into column-major tiles
//     out[row(ij) + ld * col(kl)],  row/col blocks laid out like the reference's per-quartet `buf`.
//
// Host-side work done once per (context, rank, nranks, chunk size) and cached as a JobPlan:
//   * shell pairs grouped into pair classes (la, lb, nca, ncb, padded primitive count Q), each list
//     sorted by the larger shell index so that "k <= i" is a prefix/suffix of a list;
//   * structure-of-arrays primitive tables per class (coalesced per-thread loads in kern_reg.cuh);
//   * row numbering (chunk by chunk; inside a chunk class by class in list order, so the quartets of a warp own adjacent
//     row blocks) and this rank's column numbering (kets are dealt round-robin inside every class list: static
//     sharding, no communication);
//   * chunks = ranges of the bra shell index i whose tile fits the device buffer.
// Per chunk and (T class, U class) up to two kernel launches (kets below the chunk's bra shells / the chunk's own kets):
// the register kernel if the class has one, else the cooperative kernel, else the generic block-per-quartet kernel in
// tile mode.  The same machinery runs the density-fitting job (ncenter = 3): rows = orbital pairs, kets = the
// single-shell pseudo pairs of the auxiliary shells.
#include <cstdio>
#include <cstring>
#include <cmath>
#include <vector>
#include <map>
#include <algorithm>
#include <parallel/algorithm>
#include <cuda_runtime.h>
#include "../../include/cint_b200.h"
#include "types.h"
#include "kernels.h"
#include "engine.h"
#include "rys.cuh"

int rys_tab_off(int nroots);
extern const int *engine_c2s_off();

#define CU_OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return b200_fail(CINTB200_ENODEV, "%s failed: %s", #call, cudaGetErrorString(e_)); } while (0)

#include "driver.h"

void jobplan_free(JobPlan *p)
{
    if (!p) return;
    p->arena.release();                     // every table uploaded while the plan was built
    if (p->own_out) b200_big_free(p->d_out[0]);
    b200_big_free(p->d_out[1]); b200_dfree(p->d_uprefix); b200_dfree(p->d_scratch); b200_dfree(p->d_counters);
    if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
    for (int k = 0; k < JobPlan::NS; k++) { if (p->streams[k]) cudaStreamDestroy(p->streams[k]); if (p->ev_join[k]) cudaEventDestroy(p->ev_join[k]); }
    if (p->ev_fork) cudaEventDestroy(p->ev_fork);
    for (int b = 0; b < 2; b++) {
        if (p->ev_done[b]) cudaEventDestroy(p->ev_done[b]);
        if (p->ev_copied[b]) cudaEventDestroy(p->ev_copied[b]);
    }
    if (p->ev_t0) cudaEventDestroy(p->ev_t0);
    if (p->ev_t1) cudaEventDestroy(p->ev_t1);
    digest_free(p->digest);
    for (int k = 0; k < 2; k++) if (p->graph[k].exec) cudaGraphExecDestroy(p->graph[k].exec);
    delete p;
}

void *DeviceArena::alloc(size_t bytes)
{
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes > left) {
        const size_t blk = std::max(bytes, (size_t)32 << 20);
        void *p = nullptr;
        if (b200_big_alloc(&p, blk)) return nullptr;
        char *m = (char *)malloc(blk);
        if (!m) { b200_big_free(p); return nullptr; }
        blocks.push_back(p); mirrors.push_back(m); sizes.push_back(blk); flushed.push_back(0); used.push_back(0);
        cur = (char *)p; left = blk;
    }
    void *r = cur;
    cur += bytes; left -= bytes;
    used.back() = (size_t)(cur - (char *)blocks.back());
    return r;
}
char *DeviceArena::host_of(void *dev)
{
    for (size_t k = blocks.size(); k-- > 0;)
        if ((char *)dev >= (char *)blocks[k] && (char *)dev < (char *)blocks[k] + sizes[k]) return mirrors[k] + ((char *)dev - (char *)blocks[k]);
    return nullptr;
}
int DeviceArena::flush()
{
    for (size_t k = 0; k < blocks.size(); k++) {
        if (used[k] > flushed[k]) {
            if (cudaMemcpy((char *)blocks[k] + flushed[k], mirrors[k] + flushed[k], used[k] - flushed[k], cudaMemcpyHostToDevice) != cudaSuccess) return -1;
            flushed[k] = used[k];
        }
    }
    return 0;
}
void DeviceArena::release()
{
    for (void *p : blocks) b200_big_free(p);
    for (char *m : mirrors) free(m);
    blocks.clear(); mirrors.clear(); sizes.clear(); flushed.clear(); used.clear(); cur = nullptr; left = 0;
}

static thread_local DeviceArena *g_arena = nullptr;         // set while a plan is being built: upload() allocates from it

template <class T>
static int upload(T **dst, const std::vector<T> &src)
{
    const size_t bytes = sizeof(T) * std::max<size_t>(1, src.size());
    if (g_arena) {
        *dst = (T *)g_arena->alloc(bytes);
        if (!*dst) return b200_fail(CINTB200_ENOMEM, "device arena: %zu bytes failed", bytes);
        if (!src.empty()) memcpy(g_arena->host_of(*dst), src.data(), sizeof(T) * src.size());       // sent by DeviceArena::flush()
        return 0;
    } else if (b200_dmalloc((void **)dst, bytes) != cudaSuccess)
        return b200_fail(CINTB200_ENOMEM, "cudaMalloc of %zu bytes failed", bytes);
    if (!src.empty() && cudaMemcpy(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice) != cudaSuccess)
        return b200_fail(CINTB200_ENODEV, "upload failed");
    return 0;
}

// algorithmic FLOPs per executed primitive quartet / per contracted quartet
static void model_flops(int li, int lj, int lk, int ll, int nc, double *per_prim, double *per_quartet)
{
    const int nmax = li + lj, mmax = lk + ll, n = (nmax + mmax) / 2 + 1;
    const double nf = (double)B200_NCART(li) * B200_NCART(lj) * B200_NCART(lk) * B200_NCART(ll);
    double V = 0;
    if (nmax > 0) V += 1 + 4 * (nmax - 1);
    if (mmax > 0) V += 1 + 4 * (mmax - 1);
    if (nmax > 0 && mmax > 0) V += 3 + 6 * (mmax - 1);
    V += 7.0 * mmax * std::max(nmax - 1, 0);
    double H = 0;
    const int mnij = std::min(li, lj), mxij = std::max(li, lj), mnkl = std::min(lk, ll);
    for (int a = 1; a <= mnij; a++) H += (nmax - a + 1) * (mmax + 1);
    for (int c = 1; c <= mnkl; c++) H += (mmax - c + 1) * (mnij + 1) * (mxij + 1);
    *per_prim = 40 + 58.0 * n + 23.0 * n + 3.0 * n * V + 6.0 * n * H + 3.0 * n * nf + 2.0 * nf * std::min(nc, 2);
    *per_quartet = 4.0 * nf * nc;
}

// Structure-of-arrays primitive table of a list of T pairs: rows [6 + nct][Q][NT] = aij, 1/aij, px, py, pz, kij, cc[nct]
// (consecutive THREADS read consecutive addresses); pairs are padded to Q primitives with zero-weight entries.
static void fill_tprim(const CINTOpt *c, const PairHdr &h, size_t n, size_t NT, int Q, int nct, std::vector<double> &tprim, std::vector<double> &tgeom)
{
    const size_t F = (size_t)Q * NT;
    for (int dd = 0; dd < 3; dd++) { tgeom[dd * NT + n] = h.ra[dd]; tgeom[(3 + dd) * NT + n] = h.ab[dd]; }
    for (int q = 0; q < Q; q++) {
        const size_t o = (size_t)q * NT + n;
        if (q < h.npp) {
            const PrimPair &pp = c->prims[h.pp_off + q];
            tprim[o] = pp.aij; tprim[F + o] = pp.inv_aij;
            tprim[2 * F + o] = pp.px; tprim[3 * F + o] = pp.py; tprim[4 * F + o] = pp.pz;
            tprim[5 * F + o] = pp.kij;
            for (int k = 0; k < nct; k++) tprim[(6 + k) * F + o] = c->pcoef[h.cc_off + (size_t)q * nct + k];
        } else {            // zero-weight padding primitive
            tprim[o] = 1.0; tprim[F + o] = 1.0;
            tprim[2 * F + o] = h.ra[0]; tprim[3 * F + o] = h.ra[1]; tprim[4 * F + o] = h.ra[2];
            tprim[5 * F + o] = 0.0;
            for (int k = 0; k < nct; k++) tprim[(6 + k) * F + o] = 0.0;
        }
    }
}

static int build_plan(CINTOpt *c, JobPlan *plan)
{
    struct ArenaScope { ArenaScope(DeviceArena *a) { g_arena = a; } ~ArenaScope() { if (g_arena) g_arena->flush(); g_arena = nullptr; } } arena_scope(plan->host_only ? nullptr : &plan->arena);
    const int nranks = plan->nranks, rank = plan->rank;
    const bool three = plan->ncenter == 3;
    const int nbas = three ? plan->aux0 : c->nbas;          // shells that form the bra pairs (rows)
    const size_t npair = (size_t)nbas * (nbas + 1) / 2;
    long long aux_cols = 0;                                 // 3-centre: spherical AOs of all auxiliary shells
    const int cart = plan->cart;
    auto sdim = [&](int i) { const ShellInfo &si = c->shells[i]; return (long long)(cart ? B200_NCART(si.l) : 2 * si.l + 1) * si.nctr; };
    if (three) for (int k = plan->aux0; k < c->nbas; k++) aux_cols += sdim(k);
    auto pair_dim = [&](int i, int j) { return sdim(i) * sdim(j); };
    // 1. pair classes (la, lb, nca, ncb), lists in enumeration order = sorted by the larger shell index
    std::map<std::vector<int>, int> key2class;
    plan->rowoff.resize(npair);
    plan->rows_before.assign(nbas + 1, 0);
    std::vector<long long> tcols_before(nbas + 1, 0);
    long long rows = 0, maxdim = 1;
    for (int i = 0; i < nbas; i++) {
        plan->rows_before[i] = rows;
        for (int j = 0; j <= i; j++) {
            const size_t p = (size_t)i * (i + 1) / 2 + j;
            const PairHdr &h = c->pairs[p];
            std::vector<int> key = {h.la, h.lb, h.nca, h.ncb};
            auto it = key2class.find(key);
            int ci;
            if (it == key2class.end()) {
                ci = (int)plan->classes.size();
                key2class[key] = ci;
                PairClass pc;
                pc.la = h.la; pc.lb = h.lb; pc.nca = h.nca; pc.ncb = h.ncb; pc.Q = 1;
                plan->classes.push_back(pc);
            } else ci = it->second;
            PairClass &pc = plan->classes[ci];
            pc.ids.push_back((int)p);
            pc.I.push_back(i);
            pc.npp.push_back(h.npp);
            pc.Q = std::max(pc.Q, h.npp);
            plan->rowoff[p] = rows;
            rows += pair_dim(i, j);
            maxdim = std::max(maxdim, pair_dim(i, j));
        }
    }
    plan->rows_before[nbas] = rows;
    for (int i = 0; i <= nbas; i++) tcols_before[i] = plan->rows_before[i];      // same numbering before sharding
    // 2. chunks: ranges of i; column need estimated as total/nranks + slack (exact numbering comes after the
    //    per-chunk reordering, the buffer is then sized from the exact maximum)
    const size_t cap = plan->chunk_bytes / sizeof(double);
    const long long slack = maxdim * (long long)plan->classes.size() * 2;
    // The chunk boundaries do NOT depend on the number of ranks (the estimate below is the single-rank one): every rank count
    // walks the same chunk sequence with 1/nranks of the columns, so the share of diagonal-ket work, the launch list and the
    // copied rectangle per chunk stay the same from 1 to 8 GPUs (a rank's tile is then ~chunk_bytes / nranks).
    auto est_cols = [&](int i1) { return (size_t)((three ? aux_cols : tcols_before[i1]) + slack); };
    int i0 = 0;
    while (i0 < nbas) {
        int i1 = i0 + 1;
        while (i1 < nbas) {
            size_t r = (size_t)(plan->rows_before[i1 + 1] - plan->rows_before[i0]);
            if (r * est_cols(i1 + 1) > cap) break;
            i1++;
        }
        plan->chunks.push_back({i0, i1});
        i0 = i1;
    }
    // 3. inside every chunk range, order each class list by descending primitive count: a block's first
    //    quartet then carries the block's loop bound and neighbouring threads do equal work
    for (PairClass &pc : plan->classes) {
        {
            std::vector<int> ob(pc.ids.size());
            for (size_t k = 0; k < ob.size(); k++) ob[k] = (int)k;
            std::stable_sort(ob.begin(), ob.end(), [&](int x, int y) { return pc.I[x] != pc.I[y] ? pc.I[x] < pc.I[y] : pc.npp[x] > pc.npp[y]; });
            pc.idsB.resize(ob.size()); pc.IB.resize(ob.size()); pc.nppB.resize(ob.size());
            for (size_t k = 0; k < ob.size(); k++) { pc.idsB[k] = pc.ids[ob[k]]; pc.IB[k] = pc.I[ob[k]]; pc.nppB[k] = pc.npp[ob[k]]; }
        }
        std::vector<int> order(pc.ids.size());
        for (size_t k = 0; k < order.size(); k++) order[k] = (int)k;
        for (auto &ch : plan->chunks) {
            auto b = std::lower_bound(pc.I.begin(), pc.I.end(), ch.first) - pc.I.begin();
            auto e = std::lower_bound(pc.I.begin(), pc.I.end(), ch.second) - pc.I.begin();
            std::stable_sort(order.begin() + b, order.begin() + e, [&](int x, int y) { return pc.npp[x] > pc.npp[y]; });
        }
        pc.chunk_lo.clear();
        for (auto &ch : plan->chunks) pc.chunk_lo.push_back((int)(std::lower_bound(pc.I.begin(), pc.I.end(), ch.first) - pc.I.begin()));
        pc.chunk_lo.push_back((int)pc.ids.size());
        std::vector<int> ids(order.size()), I(order.size()), npp(order.size());
        for (size_t k = 0; k < order.size(); k++) { ids[k] = pc.ids[order[k]]; I[k] = pc.I[order[k]]; npp[k] = pc.npp[order[k]]; }
        pc.ids.swap(ids); pc.I.swap(I); pc.npp.swap(npp);
    }
    // 3b. row numbering: inside every chunk the row blocks follow the class lists (class by class, in list order), so the
    //     quartets of one warp -- consecutive list entries -- own ADJACENT row blocks and the kernels' stores coalesce
    //     (a chunk still holds exactly the pairs of its bra-shell range, so rows_before[] is unchanged)
    for (size_t ch = 0; ch < plan->chunks.size(); ch++) {
        long long r = plan->rows_before[plan->chunks[ch].first];
        for (PairClass &pc : plan->classes)
            for (int k = pc.chunk_lo[ch]; k < pc.chunk_lo[ch + 1]; k++) {
                const int p = pc.ids[k], i = pc.I[k], j = p - i * (i + 1) / 2;
                plan->rowoff[p] = r;
                r += pair_dim(i, j);
            }
        if (r != plan->rows_before[plan->chunks[ch].second]) return b200_fail(CINTB200_EINVAL, "internal: row numbering of chunk %zu is inconsistent", ch);
    }
    // 4. this rank's kets: index in the final class order modulo nranks; columns numbered chunk by chunk so that
    //    the kets with K < i1 always occupy the first cols_before[i1] columns
    std::vector<long long> &colof = plan->colof;
    colof.assign(npair, -1);
    plan->cols_before.assign(nbas + 1, 0);
    long long cols = 0;
    size_t need = 1;
    if (three) {
        // kets = single-shell pseudo pairs of the auxiliary shells (pair ids npair_all + k, engine.cu:build_pairs), grouped
        // by (l, nctr); dealt round-robin to the ranks inside every class; every chunk (range of bra shells) needs all of them
        const size_t npair_all = (size_t)c->nbas * (c->nbas + 1) / 2;
        std::map<std::vector<int>, int> ukey;
        for (int k = plan->aux0; k < c->nbas; k++) {
            const ShellInfo &sk = c->shells[k];
            std::vector<int> key = {sk.l, sk.nctr};
            auto it = ukey.find(key);
            int ci;
            if (it == ukey.end()) {
                ci = (int)plan->uclasses.size();
                ukey[key] = ci;
                PairClass pc;
                pc.la = sk.l; pc.lb = 0; pc.nca = sk.nctr; pc.ncb = 1; pc.Q = 1;
                plan->uclasses.push_back(pc);
            } else ci = it->second;
            PairClass &pc = plan->uclasses[ci];
            pc.ids.push_back((int)(npair_all + k));
            pc.I.push_back(k);
            pc.npp.push_back(sk.nprim);
            pc.Q = std::max(pc.Q, sk.nprim);
        }
        plan->colof_aux.assign(c->nbas - plan->aux0, -1);
        for (PairClass &pc : plan->uclasses) {
            const size_t NU = pc.ids.size();
            std::vector<long long> ucol(NU, -1);
            std::vector<int> ustride(2 * NU, 0);
            for (size_t n = 0; n < NU; n++) {
                ustride[n] = 1;                                 // stride of the auxiliary index in columns; no second ket index
                if ((int)(n % nranks) != rank) continue;
                ucol[n] = cols;
                plan->colof_aux[pc.I[n] - plan->aux0] = cols;
                cols += (cart ? B200_NCART(pc.la) : 2 * pc.la + 1) * pc.nca;
            }
            pc.npp_prefix.assign(NU + 1, 0);
            for (size_t n = 0; n < NU; n++) pc.npp_prefix[n + 1] = pc.npp_prefix[n] + pc.npp[n];
            pc.chunk_lo.assign(plan->chunks.size() + 1, (int)NU);
            if (plan->host_only) continue;
            if (upload(&pc.d_tpair, pc.ids) || upload(&pc.d_tI, pc.I) || upload(&pc.d_ucol, ucol) || upload(&pc.d_ustride, ustride))
                return CINTB200_ENOMEM;
        }
        for (size_t ch = 0; ch < plan->chunks.size(); ch++) {
            plan->chunk_cols.push_back(cols);
            const size_t r = (size_t)(plan->rows_before[plan->chunks[ch].second] - plan->rows_before[plan->chunks[ch].first]);
            need = std::max(need, r * (size_t)std::max<long long>(cols, 1));
        }
    }
    for (size_t ch = 0; ch < plan->chunks.size() && !three; ch++) {
        for (int i = plan->chunks[ch].first; i < plan->chunks[ch].second; i++) plan->cols_before[i] = cols;   // lower bound only
        for (PairClass &pc : plan->classes)
            for (int k = pc.chunk_lo[ch]; k < pc.chunk_lo[ch + 1]; k++) {
                if (k % nranks != rank) continue;
                const int p = pc.ids[k];
                colof[p] = cols;
                const int i = pc.I[k], j = p - i * (i + 1) / 2;
                cols += pair_dim(i, j);
            }
        plan->chunk_cols.push_back(cols);
        const size_t r = (size_t)(plan->rows_before[plan->chunks[ch].second] - plan->rows_before[plan->chunks[ch].first]);
        need = std::max(need, r * (size_t)std::max<long long>(cols, 1));
    }
    plan->cols_before[nbas] = cols;
    plan->out_doubles = need;
    // 4b. geometry maps for the consumers of the tiles (digest.cu, host callbacks): bra pair + position of every row,
    //     ket pair (3-centre: auxiliary shell) + position of every column of this rank
    plan->row_pair.assign((size_t)rows, -1); plan->row_pos.assign((size_t)rows, 0);
    for (int i = 0; i < nbas; i++)
        for (int j = 0; j <= i; j++) {
            const size_t p = (size_t)i * (i + 1) / 2 + j;
            const long long n = pair_dim(i, j), r0 = plan->rowoff[p];
            for (long long r = 0; r < n; r++) { plan->row_pair[r0 + r] = (int)p; plan->row_pos[r0 + r] = (int)r; }
        }
    plan->col_pair.assign((size_t)cols, -1); plan->col_pos.assign((size_t)cols, 0);
    if (three) {
        for (int k = plan->aux0; k < c->nbas; k++) {
            const long long c0 = plan->colof_aux[k - plan->aux0];
            if (c0 < 0) continue;
            const int dk = (int)sdim(k);
            for (int r = 0; r < dk; r++) { plan->col_pair[c0 + r] = k; plan->col_pos[c0 + r] = r; }
        }
    } else {
        for (int i = 0; i < nbas; i++)
            for (int j = 0; j <= i; j++) {
                const size_t p = (size_t)i * (i + 1) / 2 + j;
                if (colof[p] < 0) continue;
                const long long n = pair_dim(i, j);
                for (long long r = 0; r < n; r++) { plan->col_pair[colof[p] + r] = (int)p; plan->col_pos[colof[p] + r] = (int)r; }
            }
    }
    // 5. device tables per class, for both orderings
    for (PairClass &pc : plan->classes) {
        const size_t NT = pc.ids.size();
        pc.npp_prefix.assign(NT + 1, 0);
        for (size_t n = 0; n < NT; n++) pc.npp_prefix[n + 1] = pc.npp_prefix[n] + pc.npp[n];
        if (plan->host_only) continue;
        const int nct = pc.nca * pc.ncb, Q = pc.Q;
        // ONE table set per class holding both orderings back to back: rows [0, NT) = ordering A (per chunk by descending primitive
        // count), rows [NT, 2 NT) = ordering B (by the larger shell index), so that a single launch can serve the kets below the
        // chunk's bra shells (all bras valid, A) and the chunk's own kets (valid bras = a suffix of B): see build_launches
        {
            const size_t N2 = 2 * NT;
            std::vector<double> tprim((size_t)(6 + nct) * Q * N2), tgeom(6 * N2);
            std::vector<long long> trow(N2), ucol(NT);
            std::vector<int> tstride(2 * N2), ustride(2 * NT), tI(N2), tids(N2), nppc(N2);
            for (size_t n = 0; n < N2; n++) {
                const bool B = n >= NT;
                const int p = B ? pc.idsB[n - NT] : pc.ids[n], i = B ? pc.IB[n - NT] : pc.I[n];
                const PairHdr &h = c->pairs[p];
                fill_tprim(c, h, n, N2, Q, nct, tprim, tgeom);
                // strides of the canonical indices inside the (i,j) block: i fastest
                const int di = (int)sdim(i);
                const bool a_is_i = (h.sh_a == i);
                tstride[n] = a_is_i ? 1 : di;  tstride[N2 + n] = a_is_i ? di : 1;
                trow[n] = plan->rowoff[p];
                tI[n] = i; tids[n] = p;
                nppc[n] = std::max(B ? pc.nppB[n - NT] : pc.npp[n], 1);       // dead pairs still run one zero-weight primitive
                if (!B) { ucol[n] = colof[p]; ustride[n] = tstride[n]; ustride[NT + n] = tstride[N2 + n]; }
            }
            if (upload(&pc.d_tprim, tprim) || upload(&pc.d_tgeom, tgeom) || upload(&pc.d_trow, trow) || upload(&pc.d_tstride, tstride) ||
                upload(&pc.d_tI, tI) || upload(&pc.d_tpair, tids) || upload(&pc.d_tnpp, nppc) || upload(&pc.d_ucol, ucol) || upload(&pc.d_ustride, ustride))
                return CINTB200_ENOMEM;
            if (!c->schwarz.empty()) {
                std::vector<double> q2(N2);
                for (size_t n = 0; n < N2; n++) q2[n] = c->schwarz[n >= NT ? pc.idsB[n - NT] : pc.ids[n]];
                if (upload(&pc.d_tq, q2)) return CINTB200_ENOMEM;
            }
        }
    }
    if (plan->host_only) return 0;
    // one tile buffer; the second one (overlap of D2H with the next chunk's kernels) is allocated on first use of a host sink
    if (b200_big_alloc((void **)&plan->d_out[0], sizeof(double) * (need + 16)))      // + slack: the bulk-copy consumers (digest.cu) round a column segment up to 16 bytes
        return b200_fail(CINTB200_ENOMEM, "cannot allocate %zu-byte tile buffer", sizeof(double) * need);
    CU_OK(cudaStreamCreateWithFlags(&plan->copy_stream, cudaStreamNonBlocking));
    for (int b = 0; b < 2; b++) {
        CU_OK(cudaEventCreateWithFlags(&plan->ev_done[b], cudaEventDisableTiming));
        CU_OK(cudaEventCreateWithFlags(&plan->ev_copied[b], cudaEventDisableTiming));
    }
    for (int k = 0; k < JobPlan::NS; k++) {
        CU_OK(cudaStreamCreateWithFlags(&plan->streams[k], cudaStreamNonBlocking));
        CU_OK(cudaEventCreateWithFlags(&plan->ev_join[k], cudaEventDisableTiming));
    }
    CU_OK(cudaEventCreateWithFlags(&plan->ev_fork, cudaEventDisableTiming));
    CU_OK(cudaEventCreate(&plan->ev_t0));
    CU_OK(cudaEventCreate(&plan->ev_t1));
    return 0;
}


static int build_launches(CINTOpt *c, JobPlan *plan)
{
    const int rank = plan->rank, nranks = plan->nranks;
    size_t scratch_need = 0;
    plan->st_quartets = plan->st_integrals = plan->st_prim = plan->st_flops = 0;
    for (size_t ch = 0; ch < plan->chunks.size(); ch++) {
        const int i0 = plan->chunks[ch].first, i1 = plan->chunks[ch].second;
        const long long row0 = plan->rows_before[i0];
        const long long ld = plan->rows_before[i1] - row0;
        if (ld == 0 || plan->chunk_cols[ch] == 0) continue;
        for (size_t ct = 0; ct < plan->classes.size(); ct++) {
            PairClass &T = plan->classes[ct];
            const int t_begin = T.chunk_lo[ch], t_end = T.chunk_lo[ch + 1];
            if (t_end <= t_begin) continue;
            // per bra shell index inside this chunk: number of T pairs / primitives with I >= i
            std::vector<double> cnt_ge(i1 - i0 + 1, 0.0), npp_ge(i1 - i0 + 1, 0.0);
            for (int t = t_begin; t < t_end; t++) { cnt_ge[T.I[t] - i0] += 1; npp_ge[T.I[t] - i0] += T.npp[t]; }
            for (int k = i1 - i0 - 1; k >= 0; k--) { cnt_ge[k] += cnt_ge[k + 1]; npp_ge[k] += npp_ge[k + 1]; }
            std::vector<PairClass> &UC = (plan->ncenter == 3) ? plan->uclasses : plan->classes;
            for (size_t cu = 0; cu < UC.size(); cu++) {
              PairClass &U = UC[cu];
              // part 0: kets below the chunk's shell range (every bra of the chunk is valid; bras sorted by primitive count: ordering A)
              // part 1: kets inside it (valid bras = those with I >= K: a suffix of the shell-sorted ordering B)
              // part 2: both in ONE launch (the kernel picks the ordering per ket): half the launches; CINTB200_MERGE=0 keeps them apart
              // 3-centre jobs: every auxiliary ket meets every bra pair of the chunk -> part 0 only, all kets
              static const bool merge = !(getenv("CINTB200_MERGE") && !atoi(getenv("CINTB200_MERGE")));
              const int part_first = (plan->ncenter == 3) ? 0 : merge ? 2 : 0, part_last = (plan->ncenter == 3) ? 0 : merge ? 2 : 1;
              for (int part = part_first; part <= part_last; part++) {
                const int u_lo = part == 1 ? U.chunk_lo[ch] : 0, u_hi = part == 0 ? U.chunk_lo[ch] : U.chunk_lo[ch + 1];
                // this rank's kets in [u_lo, u_hi): indices congruent to rank modulo nranks
                int u_first = u_lo + ((rank - u_lo) % nranks + nranks) % nranks;
                if (u_first >= u_hi) continue;
                const int nu_mine = (u_hi - u_first + nranks - 1) / nranks;
                LaunchRec L;
                memset(&L, 0, sizeof L);
                TileParams &P = L.P;
                P.tprim = T.d_tprim; P.tgeom = T.d_tgeom; P.trow = T.d_trow; P.tstride = T.d_tstride;
                P.tI = T.d_tI; P.tpair = T.d_tpair; P.tnpp = T.d_tnpp; P.tq = T.d_tq;
                const int NT1 = (int)T.ids.size(), tB = (plan->ncenter == 3) ? 0 : NT1;       // rect / 3-centre plans hold ordering A only
                P.NT = plan->rect ? NT1 : 2 * NT1; P.NTs = P.NT; P.Q = T.Q; P.nca_t = T.nca;
                P.t_begin = t_begin + (part == 1 ? tB : 0); P.t_end = t_end + (part == 1 ? tB : 0);
                P.tB = part == 2 ? tB : 0; P.tri_i0 = i0;
                P.upair = U.d_tpair; P.uK = U.d_tI; P.ucol = U.d_ucol; P.ustride = U.d_ustride;
                P.NU = nu_mine; P.NU_all = (int)U.ids.size(); P.u_step = nranks; P.u_first = u_first; P.nca_u = U.nca; P.umax = std::max(1, U.Q);
                P.tri = part;
                P.uq = U.d_tq; P.schwarz_thr = (T.d_tq && U.d_tq) ? c->schwarz_thr : 0.0;
                P.row0 = row0; P.ld = ld;
                P.pairs = c->d_pairs; P.prims = c->d_prims; P.pcoef = c->d_pcoef;
                L.chunk = (int)ch;
                L.nroots = (T.la + T.lb + U.la + U.lb) / 2 + 1;
                L.ncu = U.nca * U.ncb;
                P.rys = c->d_rys + rys_tab_off(L.nroots);
                double q_here = 0, prim_here = 0;
                for (int j = 0; j < nu_mine; j++) {
                    const int u = u_first + nranks * j;
                    const int kk = (plan->ncenter == 3) ? 0 : std::max(U.I[u], i0) - i0;   // K < i0: every T pair of the chunk is valid
                    q_here += cnt_ge[kk];
                    prim_here += (double)U.npp[u] * npp_ge[kk];
                }
                if (q_here == 0) continue;
                auto ld_ = [&](int l) { return plan->cart ? B200_NCART(l) : 2 * l + 1; };
                const double blk = (double)ld_(T.la) * ld_(T.lb) * T.nca * T.ncb * ld_(U.la) * ld_(U.lb) * U.nca * U.ncb;
                double fp, fq;
                model_flops(T.la, T.lb, U.la, U.lb, T.nca * T.ncb * U.nca * U.ncb, &fp, &fq);
                plan->st_quartets += q_here; plan->st_integrals += q_here * blk; plan->st_prim += prim_here;
                plan->st_flops += prim_here * fp + q_here * fq;
                L.ntasks = (long long)(t_end - t_begin) * nu_mine;      // generic: rectangle, invalid quartets skipped in-kernel
                L.key[0] = T.la; L.key[1] = T.lb; L.key[2] = U.la; L.key[3] = U.lb; L.key[4] = T.nca * T.ncb; L.key[5] = U.nca * U.ncb;
                L.quartets = q_here; L.prim = prim_here; L.flops = prim_here * fp + q_here * fq; L.integrals = q_here * blk; L.part = part;
                // range-separated operators run on the RS instantiations of the same kernels (TileParams::rs_*)
                const bool generic_only = c->force_generic;
                const int rs = c->omega != 0;
                P.rs_w2 = c->omega * c->omega; P.rs_sign = c->omega;      /* sign * |omega| */ P.rs_pass0 = c->omega > 0 ? 1 : 0;
                L.fn = generic_only ? nullptr : reg_kernel_lookup(T.la, T.lb, U.la, U.lb, T.nca * T.ncb, U.nca * U.ncb, rs, plan->cart);
                if (!L.fn && !generic_only) {
                    L.fn = coop_kernel_lookup(T.la, T.lb, U.la, U.lb, T.nca * T.ncb, U.nca * U.ncb, &L.ci, rs, plan->cart);
                    L.coop = L.fn != nullptr;
                }
                // deeply contracted kets (all-electron heavy-element bases) can push the staged ket primitives past the
                // shared memory a block may have: such classes go to the generic kernel instead of failing at launch
                if (L.fn && !plan->host_only && (L.coop ? coop_kernel_smem(L.ci, L.ncu, P.umax) : reg_kernel_smem(L.fn, L.nroots, L.ncu, P.umax)) > (size_t)tile_smem_limit()) {
                    L.fn = nullptr; L.coop = 0;
                }
                if (L.fn && !L.coop && REG_FAST_RYS && L.nroots <= RYS_FNMAX && L.nroots <= REG_FAST_NMAX)
                    P.rys = c->d_rys_fast + rys_fast_off(L.nroots);   // register kernels of low order read the degree-6 tables
                if (L.fn) {
                    const int qpb = L.coop ? 32 / L.ci.fs : 32;       // bras per work item (one warp)
                    L.gx = (t_end - t_begin + qpb - 1) / qpb;
                    L.gy = nu_mine;
                    L.P.gx = L.gx;
                    {
                        const long long items = (long long)L.gx * nu_mine;
                        L.P.batch = (int)std::max<long long>(1, std::min<long long>(16, items / (148 * 32 * 4)));
                    }
                    plan->launches.push_back(L);
                } else {
                    if (generic_plan(&L.GC, &L.GL, T.la, T.lb, U.la, U.lb, T.nca * T.ncb, U.nca * U.ncb, plan->cart, L.ntasks, engine_c2s_off(), c->omega < 0))
                        return b200_fail(CINTB200_ENOSUP, "class (%d%d|%d%d) exceeds this build's limits", T.la, T.lb, U.la, U.lb);
                    scratch_need = std::max(scratch_need, L.GC.scratch_per_block * (size_t)L.GL.grid);
                    plan->launches.push_back(L);
                }
              }
            }
        }
    }
    // longest kernels first inside every chunk: the short ones then fill the tail on the other streams
    std::stable_sort(plan->launches.begin(), plan->launches.end(), [](const LaunchRec &a, const LaunchRec &b) {
        return a.chunk != b.chunk ? a.chunk < b.chunk : a.flops > b.flops; });
    if (!plan->host_only) {
        if (b200_dmalloc((void **)&plan->d_counters, sizeof(unsigned int) * std::max<size_t>(1, plan->launches.size())) != cudaSuccess)
            return b200_fail(CINTB200_ENOMEM, "cannot allocate launch counters");
        for (size_t k = 0; k < plan->launches.size(); k++) plan->launches[k].P.counter = plan->d_counters + k;
    }
    if (!plan->host_only && scratch_need && b200_dmalloc((void **)&plan->d_scratch, sizeof(double) * scratch_need) != cudaSuccess)
        return b200_fail(CINTB200_ENOMEM, "cannot allocate generic-kernel scratch");
    return 0;
}

// How the finished tiles leave the device (or do not): host sinks, per-tile callback, device-side consumers
struct TileSink {
    double *const *sinks = nullptr;         // ring of pinned host buffers (each >= the largest tile); NULL: tiles stay on the device
    int nsinks = 0;
    cintb200_tile_fn fn = nullptr;          // called on the calling thread when a tile has arrived in its sink
    void *user = nullptr;
    DigestJob job;                          // device-side consumers (digest.cu)
    int cart = 0;                           // Cartesian output
    const double *dm_dev = nullptr;         // J/K: density matrix (device), results (device)
    double *vj_dev = nullptr, *vk_dev = nullptr;
};

static int execute_plan(cintb200_ctx *c, JobPlan *plan, int ncenter, const TileSink &sink, double *stats);

// ncenter = 4: every unique quartet of int2e_sph; ncenter = 3: every triple (ij|k), i >= j < aux0 <= k, of int3c2e_sph
static int run_job(cintb200_ctx *c, int ncenter, int aux0, int rank, int nranks, size_t chunk_bytes,
                   const TileSink &sink, double *stats)
{
    if (!c || c->magic != B200_CTX_MAGIC) return b200_fail(CINTB200_EINVAL, "invalid context");
    if (nranks < 1 || rank < 0 || rank >= nranks) return b200_fail(CINTB200_EINVAL, "bad rank %d of %d", rank, nranks);
    if (ncenter == 3 && (aux0 < 1 || aux0 >= c->nbas))
        return b200_fail(CINTB200_EINVAL, "first auxiliary shell %d outside 1..%d", aux0, c->nbas - 1);
    if (sink.sinks) {
        if (sink.nsinks < 1) return b200_fail(CINTB200_EINVAL, "host sinks given but nsinks = %d", sink.nsinks);
        for (int k = 0; k < sink.nsinks; k++) if (!sink.sinks[k]) return b200_fail(CINTB200_EINVAL, "host sink %d is NULL", k);
    } else if (sink.fn) return b200_fail(CINTB200_EINVAL, "a tile callback needs at least one host sink");
    double tph = b200_now();
    if (ncenter == 4 && c->schwarz_thr > 0 && c->omega == 0 && !c->force_generic && ctx_compute_schwarz(c)) return CINTB200_ENODEV;
    b200_phase("job: schwarz bounds", tph);
    std::lock_guard<std::mutex> lock(c->mtx);
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    CU_OK(cudaSetDevice(c->device));
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev_dev};
    if (chunk_bytes == 0) chunk_bytes = (size_t)16 << 30;
    JobPlan *plan = c->plan;
    if (!plan || plan->rect || plan->rank != rank || plan->nranks != nranks || plan->chunk_bytes != chunk_bytes || plan->force_generic != c->force_generic
        || plan->schwarz_thr != c->schwarz_thr || plan->ncenter != ncenter || plan->aux0 != (ncenter == 3 ? aux0 : 0) || plan->cart != sink.cart) {
        if (plan) { cudaDeviceSynchronize(); jobplan_free(plan); c->plan = nullptr; }
        plan = new JobPlan();
        plan->ncenter = ncenter; plan->aux0 = (ncenter == 3) ? aux0 : 0; plan->cart = sink.cart;
        plan->rank = rank; plan->nranks = nranks; plan->chunk_bytes = chunk_bytes; plan->force_generic = c->force_generic; plan->schwarz_thr = c->schwarz_thr;
        tph = b200_now();
        int rc = build_plan(c, plan);
        b200_phase("job: build_plan", tph);
        tph = b200_now();
        if (!rc) rc = build_launches(c, plan);
        b200_phase("job: build_launches", tph);
        if (rc) { jobplan_free(plan); return rc; }
        c->plan = plan;
    }
    if (sink.sinks && !plan->d_out[1] && b200_big_alloc((void **)&plan->d_out[1], sizeof(double) * (plan->out_doubles + 16)))
        return b200_fail(CINTB200_ENOMEM, "cannot allocate the second %zu-byte tile buffer", sizeof(double) * plan->out_doubles);
    if (sink.sinks && plan->out_doubles * sizeof(double) > chunk_bytes)
        return b200_fail(CINTB200_EINVAL, "host sinks: the largest tile (one bra shell x all of this rank's kets) needs %zu bytes, "
                         "chunk_bytes = %zu is too small", plan->out_doubles * sizeof(double), chunk_bytes);
    tph = b200_now();
    const int rc_exec = execute_plan(c, plan, ncenter, sink, stats);
    b200_phase("job: execute", tph);
    return rc_exec;
}

// Launch every kernel of a plan (all chunks); finished tiles go to the device-side consumers and / or to the host sinks.
// With host sinks the tiles alternate between two device buffers and chunk k is copied to sinks[k % nsinks] while the kernels
// of chunk k+1 run; the callback for chunk k is made on the calling thread once that copy has completed and before the sink
// is reused, so a consumer sees every tile.  Caller holds c->mtx.
static int execute_plan(cintb200_ctx *c, JobPlan *plan, int ncenter, const TileSink &sink, double *stats)
{
    EngineParams EP;
    EP.pairs = c->d_pairs; EP.prims = c->d_prims; EP.pcoef = c->d_pcoef; EP.rys_coef = c->d_rys; EP.c2s = c->d_c2s;
    EP.expcutoff = (ncenter == 3) ? c->expcutoff3 : c->expcutoff4; EP.omega = c->omega; EP.cart = plan->cart;

    double d2h = 0;
    long long nlaunch = 0, reg_launches = 0;
    const bool prof = c->profile != 0;
    const bool host = sink.sinks != nullptr;
    cudaStream_t st = c->stream;
    const int NS = prof ? 1 : JobPlan::NS;
    auto fork = [&]() -> int {                 // side streams start after everything queued on the main stream
        if (NS == 1) return 0;
        CU_OK(cudaEventRecord(plan->ev_fork, st));
        for (int k = 0; k < NS; k++) CU_OK(cudaStreamWaitEvent(plan->streams[k], plan->ev_fork, 0));
        return 0;
    };
    auto join = [&]() -> int {                 // main stream continues after all side streams drained
        if (NS == 1) return 0;
        for (int k = 0; k < NS; k++) {
            CU_OK(cudaEventRecord(plan->ev_join[k], plan->streams[k]));
            CU_OK(cudaStreamWaitEvent(st, plan->ev_join[k], 0));
        }
        return 0;
    };
    auto tile_bytes = [&](int ch) {
        const int i0 = plan->chunks[ch].first, i1 = plan->chunks[ch].second;
        return sizeof(double) * (size_t)(plan->rows_before[i1] - plan->rows_before[i0]) * (size_t)plan->chunk_cols[ch];
    };
    // chunk done on the device: consumers, then the copy to host sink number `slot`
    auto finish_chunk = [&](int ch, int b, int slot) -> int {
        if (join()) return CINTB200_ENODEV;
        if (digest_tile(c, plan, sink.job, ch, plan->d_out[b], st)) return CINTB200_ENODEV;
        CU_OK(cudaEventRecord(plan->ev_done[b], st));
        if (host) {
            const size_t bytes = tile_bytes(ch);
            CU_OK(cudaStreamWaitEvent(plan->copy_stream, plan->ev_done[b], 0));
            CU_OK(cudaMemcpyAsync(sink.sinks[slot], plan->d_out[b], bytes, cudaMemcpyDeviceToHost, plan->copy_stream));
            CU_OK(cudaEventRecord(plan->ev_copied[b], plan->copy_stream));
            d2h += (double)bytes;
        }
        return 0;
    };
    // hand a tile that has been queued for copying to the caller
    auto deliver = [&](int ch, int b, int slot) -> int {
        if (!sink.fn) return 0;
        CU_OK(cudaEventSynchronize(plan->ev_copied[b]));
        cintb200_tile t;
        memset(&t, 0, sizeof t);
        t.chunk = ch; t.nchunks = (int)plan->chunks.size(); t.rank = plan->rank; t.nranks = plan->nranks;
        t.i0 = plan->chunks[ch].first; t.i1 = plan->chunks[ch].second;
        t.row0 = plan->rows_before[t.i0]; t.nrows = plan->rows_before[t.i1] - t.row0; t.ncols = plan->chunk_cols[ch];
        t.ncols_below = (plan->ncenter == 3) ? t.ncols : plan->cols_before[t.i0];
        if (sink.fn(sink.user, &t, sink.sinks[slot]) != 0) return b200_fail(CINTB200_EINVAL, "tile callback asked to stop at chunk %d", ch);
        return 0;
    };
    std::vector<cudaEvent_t> pev;
    if (prof) {
        pev.resize(plan->launches.size() + 1);
        for (auto &e : pev) CU_OK(cudaEventCreate(&e));
    }
    size_t li = 0;
    int pend_ch = -1, pend_buf = 0, pend_slot = 0;
    // everything that goes to the streams between the two timing events (capturable into a CUDA graph, see below)
    auto enqueue = [&]() -> int {
    int group = -1, cur_chunk = -1, buf = 0;
    CU_OK(cudaMemsetAsync(plan->d_counters, 0, sizeof(unsigned int) * std::max<size_t>(1, plan->launches.size()), st));
    if (digest_begin(c, plan, sink.job, sink.dm_dev, st)) return CINTB200_ENODEV;
    const size_t nl = plan->launches.size();
    for (size_t k = 0; k <= nl; k++) {
        const bool last = (k == nl);
        if (last || plan->launches[k].chunk != cur_chunk) {
            if (cur_chunk >= 0) {
                // the kernels of cur_chunk are queued: while they run, hand over the previous tile, then queue this one's copy
                if (pend_ch >= 0) { if (deliver(pend_ch, pend_buf, pend_slot)) return CINTB200_EINVAL; pend_ch = -1; }
                const int slot = host ? group % sink.nsinks : 0;
                if (finish_chunk(cur_chunk, buf, slot)) return CINTB200_ENODEV;
                if (host) { pend_ch = cur_chunk; pend_buf = buf; pend_slot = slot; }
            }
            if (last) break;
            cur_chunk = plan->launches[k].chunk;
            group++;
            buf = host ? (group & 1) : 0;
            if (host) {
                CU_OK(cudaStreamWaitEvent(st, plan->ev_copied[buf], 0));   // device buffer drained?
                // entries of the chunk's own kets with K > I are not written by the kernels: hand out zeros, not stale values
                const int i0 = plan->chunks[cur_chunk].first, i1 = plan->chunks[cur_chunk].second;
                const long long ld = plan->rows_before[i1] - plan->rows_before[i0];
                const long long cb = (plan->ncenter == 3) ? plan->chunk_cols[cur_chunk] : plan->cols_before[i0];
                if (!plan->rect && plan->chunk_cols[cur_chunk] > cb)
                    CU_OK(cudaMemsetAsync(plan->d_out[buf] + ld * cb, 0, sizeof(double) * (size_t)ld * (size_t)(plan->chunk_cols[cur_chunk] - cb), st));
            }
            if (fork()) return CINTB200_ENODEV;
        }
        LaunchRec &L = plan->launches[k];
        L.P.out = plan->d_out[buf];
        cudaStream_t ls = (NS == 1) ? st : plan->streams[nlaunch % NS];
        if (prof) CU_OK(cudaEventRecord(pev[li++], st));
        if (L.fn) {
            if (L.coop ? coop_kernel_launch(L.fn, L.ci, L.ncu, L.P, L.gx, L.gy, ls)
                       : reg_kernel_launch(L.fn, L.nroots, L.ncu, L.P, L.gx, L.gy, ls))
                return b200_fail(CINTB200_ENODEV, "register kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            reg_launches++;
        } else {
            L.GC.scratch = plan->d_scratch;
            ls = st;            // generic launches share one scratch area: keep them ordered on the main stream
            if (generic_launch(EP, L.GC, L.GL, nullptr, L.ntasks, plan->d_out[buf], nullptr, nullptr, ls, &L.P, nullptr))
                return b200_fail(CINTB200_ENODEV, "generic kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        }
        nlaunch++;
    }
    if (prof) CU_OK(cudaEventRecord(pev[li], st));
    if (digest_end(c, plan, sink.job, sink.vj_dev, sink.vk_dev, st)) return CINTB200_ENODEV;
    return 0;
    };      // enqueue
    // Device-resident runs replay a CUDA graph: the launch list of a plan is static (1573 launches for C60, 120-1200 for the
    // C2H6 bases), so the second run with the same consumers is captured -- fork / join over the side streams included -- and
    // every later run is ONE graph launch instead of hundreds of kernel launches (what bounds the small-molecule jobs and the
    // per-rank work at 8 GPUs).  Host-sink runs (callbacks between chunks), profiled runs and J/K runs (caller-owned
    // pointers in the kernel arguments) are enqueued directly.
    static const bool graphs_on = !(getenv("CINTB200_NO_GRAPH") && atoi(getenv("CINTB200_NO_GRAPH")));
    JobPlan::GraphSlot &gs = plan->graph[sink.job.checksums ? 1 : 0];
    const bool graphable = graphs_on && !host && !prof && !sink.job.jk;
    CU_OK(cudaEventRecord(plan->ev_t0, st));
    if (graphable && gs.exec) {
        CU_OK(cudaGraphLaunch(gs.exec, st));
        nlaunch = gs.nlaunch; reg_launches = gs.reg_launches;
        if (sink.job.checksums) digest_mark_rowsums(plan);
    } else if (graphable && gs.warm) {
        CU_OK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        const int rc_enq = enqueue();
        cudaGraph_t g = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(st, &g);
        if (rc_enq || ce != cudaSuccess || !g) {
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
            return rc_enq ? rc_enq : b200_fail(CINTB200_ENODEV, "CUDA graph capture of the job failed: %s", cudaGetErrorString(ce));
        }
        const cudaError_t ie = cudaGraphInstantiate(&gs.exec, g, 0);
        cudaGraphDestroy(g);
        if (ie != cudaSuccess) { gs.exec = nullptr; return b200_fail(CINTB200_ENODEV, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ie)); }
        gs.nlaunch = nlaunch; gs.reg_launches = reg_launches;
        CU_OK(cudaGraphLaunch(gs.exec, st));
    } else {
        const int rc_enq = enqueue();
        if (rc_enq) return rc_enq;
        if (graphable) gs.warm = 1;
    }
    CU_OK(cudaEventRecord(plan->ev_t1, st));
    if (pend_ch >= 0 && deliver(pend_ch, pend_buf, pend_slot)) return CINTB200_EINVAL;
    CU_OK(cudaStreamSynchronize(st));
    if (host) CU_OK(cudaStreamSynchronize(plan->copy_stream));
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) return b200_fail(CINTB200_ENODEV, "kernel execution failed: %s", cudaGetErrorString(le));
    float ms = 0;
    CU_OK(cudaEventElapsedTime(&ms, plan->ev_t0, plan->ev_t1));
    if (prof) {
        // aggregate launch durations by kernel class: rows of 12 doubles
        std::map<std::vector<int>, std::vector<double>> agg;
        for (size_t k = 0; k < plan->launches.size(); k++) {
            const LaunchRec &L = plan->launches[k];
            float dt = 0;
            cudaEventElapsedTime(&dt, pev[k], pev[k + 1]);
            std::vector<int> key(L.key, L.key + 6);
            key.push_back(L.fn ? (L.coop ? 2 : 1) : 0);
            std::vector<double> &a = agg[key];
            if (a.empty()) a.assign(5, 0.0);
            a[0] += dt; a[1] += L.quartets; a[2] += L.prim; a[3] += L.flops; a[4] += 1;
        }
        c->profile_rows.clear();
        for (auto &kv : agg) {
            for (int v : kv.first) c->profile_rows.push_back(v);
            for (double v : kv.second) c->profile_rows.push_back(v);
        }
        for (auto &e : pev) cudaEventDestroy(e);
    }
    if (stats) {
        double total = 0;                   // sum of all integrals: available when the checksum consumer ran
        if (sink.job.checksums && digest_fetch_checksums(c, plan, nullptr, nullptr, nullptr, &total) < 0) return CINTB200_ENODEV;
        stats[0] = plan->st_quartets; stats[1] = plan->st_integrals; stats[2] = plan->st_prim; stats[3] = total;
        stats[4] = (double)nlaunch; stats[5] = d2h; stats[6] = plan->st_flops; stats[7] = ms;
        stats[8] = (double)reg_launches; stats[9] = (double)plan->chunks.size();
        stats[10] = (double)plan->out_doubles * 8; stats[11] = (double)plan->classes.size();
    }
    c->launches += nlaunch;
    return 0;
}

extern "C" int cintb200_int2e_sph_all_unique(cintb200_ctx *c, int rank, int nranks, size_t chunk_bytes,
                                             double *host_sink, double *stats)
{
    TileSink s;
    double *one[1] = {host_sink};
    if (host_sink) { s.sinks = one; s.nsinks = 1; }
    s.job.checksums = (c && c->magic == B200_CTX_MAGIC) ? c->checksums : 0;
    return run_job(c, 4, 0, rank, nranks, chunk_bytes, s, stats);
}

extern "C" int cintb200_int3c2e_sph_all(cintb200_ctx *c, int aux_shell0, int rank, int nranks, size_t chunk_bytes,
                                        double *host_sink, double *stats)
{
    TileSink s;
    double *one[1] = {host_sink};
    if (host_sink) { s.sinks = one; s.nsinks = 1; }
    s.job.checksums = (c && c->magic == B200_CTX_MAGIC) ? c->checksums : 0;
    return run_job(c, 3, aux_shell0, rank, nranks, chunk_bytes, s, stats);
}

// int2e_cart over the same loop (north star: int2e_cart is part of the hot path): Cartesian block dimensions everywhere
extern "C" int cintb200_int2e_cart_all_unique(cintb200_ctx *c, int rank, int nranks, size_t chunk_bytes, double *host_sink, double *stats)
{
    TileSink s;
    double *one[1] = {host_sink};
    if (host_sink) { s.sinks = one; s.nsinks = 1; }
    s.cart = 1;
    s.job.checksums = (c && c->magic == B200_CTX_MAGIC) ? c->checksums : 0;
    return run_job(c, 4, 0, rank, nranks, chunk_bytes, s, stats);
}
extern "C" int cintb200_int2e_cart_all_unique_tiles(cintb200_ctx *c, int rank, int nranks, size_t chunk_bytes, double *const *sinks, int nsinks,
                                                    cintb200_tile_fn fn, void *user, double *stats)
{
    if (!sinks || nsinks < 1) return b200_fail(CINTB200_EINVAL, "cintb200_int2e_cart_all_unique_tiles needs at least one host sink");
    TileSink s;
    s.sinks = sinks; s.nsinks = nsinks; s.fn = fn; s.user = user; s.cart = 1;
    s.job.checksums = (c && c->magic == B200_CTX_MAGIC) ? c->checksums : 0;
    return run_job(c, 4, 0, rank, nranks, chunk_bytes, s, stats);
}

// Same jobs with every tile handed to the caller: ring of pinned sinks + callback (include/cint_b200.h)
extern "C" int cintb200_int2e_sph_all_unique_tiles(cintb200_ctx *c, int rank, int nranks, size_t chunk_bytes, double *const *sinks, int nsinks,
                                                   cintb200_tile_fn fn, void *user, double *stats)
{
    if (!sinks || nsinks < 1) return b200_fail(CINTB200_EINVAL, "cintb200_int2e_sph_all_unique_tiles needs at least one host sink");
    TileSink s;
    s.sinks = sinks; s.nsinks = nsinks; s.fn = fn; s.user = user;
    s.job.checksums = (c && c->magic == B200_CTX_MAGIC) ? c->checksums : 0;
    return run_job(c, 4, 0, rank, nranks, chunk_bytes, s, stats);
}

extern "C" int cintb200_int3c2e_sph_all_tiles(cintb200_ctx *c, int aux_shell0, int rank, int nranks, size_t chunk_bytes, double *const *sinks,
                                              int nsinks, cintb200_tile_fn fn, void *user, double *stats)
{
    if (!sinks || nsinks < 1) return b200_fail(CINTB200_EINVAL, "cintb200_int3c2e_sph_all_tiles needs at least one host sink");
    TileSink s;
    s.sinks = sinks; s.nsinks = nsinks; s.fn = fn; s.user = user;
    s.job.checksums = (c && c->magic == B200_CTX_MAGIC) ? c->checksums : 0;
    return run_job(c, 3, aux_shell0, rank, nranks, chunk_bytes, s, stats);
}

// Coulomb / exchange matrices digested on the device from the tiles of the whole job (digest.cu); dm, vj, vk are nao x nao
extern "C" int cintb200_int2e_sph_jk(cintb200_ctx *c, int rank, int nranks, size_t chunk_bytes, const double *dm, double *vj, double *vk,
                                     int on_device, double *stats)
{
    if (!c || c->magic != B200_CTX_MAGIC) return b200_fail(CINTB200_EINVAL, "invalid context");
    if (!dm || (!vj && !vk)) return b200_fail(CINTB200_EINVAL, "cintb200_int2e_sph_jk: dm and at least one of vj, vk are required");
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    CU_OK(cudaSetDevice(c->device));
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev_dev};
    const size_t n2 = (size_t)c->nao_sph * c->nao_sph;
    double *d_io = nullptr;
    TileSink s;
    s.job.jk = 1; s.job.want_k = vk != nullptr; s.job.checksums = c->checksums;
    if (on_device) { s.dm_dev = dm; s.vj_dev = vj; s.vk_dev = vk; }
    else {
        if (b200_dmalloc((void **)&d_io, sizeof(double) * 3 * n2) != cudaSuccess) return b200_fail(CINTB200_ENOMEM, "J/K: cannot allocate device matrices");
        if (cudaMemcpy(d_io, dm, sizeof(double) * n2, cudaMemcpyHostToDevice) != cudaSuccess) { b200_dfree(d_io); return b200_fail(CINTB200_ENODEV, "J/K: upload of dm failed"); }
        s.dm_dev = d_io; s.vj_dev = vj ? d_io + n2 : nullptr; s.vk_dev = vk ? d_io + 2 * n2 : nullptr;
    }
    int rc = run_job(c, 4, 0, rank, nranks, chunk_bytes, s, stats);
    if (!rc && !on_device) {
        if (vj && cudaMemcpy(vj, d_io + n2, sizeof(double) * n2, cudaMemcpyDeviceToHost) != cudaSuccess) rc = b200_fail(CINTB200_ENODEV, "J/K: download failed");
        if (vk && cudaMemcpy(vk, d_io + 2 * n2, sizeof(double) * n2, cudaMemcpyDeviceToHost) != cudaSuccess) rc = b200_fail(CINTB200_ENODEV, "J/K: download failed");
    }
    b200_dfree(d_io);
    return rc;
}

// ---- job geometry and checksums for the consumers of the tiles ----
extern "C" int cintb200_set_checksums(cintb200_ctx *c, int on)
{
    if (!c || c->magic != B200_CTX_MAGIC) return b200_fail(CINTB200_EINVAL, "invalid context");
    c->checksums = on != 0;
    return 0;
}

extern "C" int cintb200_job_checksums(cintb200_ctx *c, double *S, double *A, double *F)
{
    if (!c || c->magic != B200_CTX_MAGIC || !c->plan) return b200_fail(CINTB200_EINVAL, "no whole-job run to take checksums from");
    std::lock_guard<std::mutex> lock(c->mtx);
    return digest_fetch_checksums(c, c->plan, S, A, F, nullptr);
}

// geom: {i0, i1, row0, nrows, ncols, nchunks, total rows, this rank's total columns, columns of kets below the chunk's bra shells}
extern "C" int cintb200_job_geometry(cintb200_ctx *c, int chunk, long long *geom)
{
    if (!c || c->magic != B200_CTX_MAGIC || !c->plan || c->plan->rect) return b200_fail(CINTB200_EINVAL, "no whole-job plan (run a whole-job call first)");
    JobPlan *plan = c->plan;
    if (chunk < 0 || chunk >= (int)plan->chunks.size() || !geom) return b200_fail(CINTB200_EINVAL, "chunk %d out of range", chunk);
    const int i0 = plan->chunks[chunk].first, i1 = plan->chunks[chunk].second;
    geom[0] = i0; geom[1] = i1; geom[2] = plan->rows_before[i0]; geom[3] = plan->rows_before[i1] - plan->rows_before[i0];
    geom[4] = plan->chunk_cols[chunk]; geom[5] = (long long)plan->chunks.size();
    geom[6] = (long long)plan->row_pair.size(); geom[7] = (long long)plan->col_pair.size();
    geom[8] = (plan->ncenter == 3) ? plan->chunk_cols[chunk] : plan->cols_before[i0];
    return 0;
}

static inline void pair_to_shells(int p, int *i, int *j)
{
    int ii = (int)((sqrt(8.0 * p + 1.0) - 1.0) / 2.0);
    while ((long long)(ii + 1) * (ii + 2) / 2 <= p) ii++;
    while ((long long)ii * (ii + 1) / 2 > p) ii--;
    *i = ii; *j = p - ii * (ii + 1) / 2;
}

// rows of chunk `chunk` (nrows entries each): bra shells i >= j of the row and its position mi + di * mj inside the (i,j) block
extern "C" int cintb200_job_row_map(cintb200_ctx *c, int chunk, int *sh_i, int *sh_j, int *pos)
{
    long long g[9];
    if (cintb200_job_geometry(c, chunk, g)) return CINTB200_EINVAL;
    JobPlan *plan = c->plan;
    for (long long r = 0; r < g[3]; r++) {
        int i, j;
        pair_to_shells(plan->row_pair[g[2] + r], &i, &j);
        if (sh_i) sh_i[r] = i;
        if (sh_j) sh_j[r] = j;
        if (pos) pos[r] = plan->row_pos[g[2] + r];
    }
    return 0;
}

// columns of chunk `chunk` (ncols entries each): ket shells k >= l (3-centre jobs: auxiliary shell k, l = -1) and position mk + dk * ml
extern "C" int cintb200_job_col_map(cintb200_ctx *c, int chunk, int *sh_k, int *sh_l, int *pos)
{
    long long g[9];
    if (cintb200_job_geometry(c, chunk, g)) return CINTB200_EINVAL;
    JobPlan *plan = c->plan;
    for (long long q = 0; q < g[4]; q++) {
        const int p = plan->col_pair[q];
        if (plan->ncenter == 3) { if (sh_k) sh_k[q] = p; if (sh_l) sh_l[q] = -1; }
        else {
            int k, l;
            pair_to_shells(p, &k, &l);
            if (sh_k) sh_k[q] = k;
            if (sh_l) sh_l[q] = l;
        }
        if (pos) pos[q] = plan->col_pos[q];
    }
    return 0;
}

// ------------------------------------------------------------------ list mode on the tile kernels
// Arbitrary lists of shell tuples (cintb200_int2e_batch & co.) used to run on the block-per-tuple generic kernel only
// (C60: 1.2e8 integrals/s, slower than the reference on 16 cores).  Here every group of tuples with the same bra class and
// ket class that has a specialised kernel is turned into explicit work items for it: the tuples are sorted by ket, each ket
// owns a run of T OCCURRENCES (output offset + strides of that tuple, and the row of the bra pair in a per-class pair table
// shared by all calls), and an item is {ket, first occurrence, count <= 32}.  Groups without a kernel stay with the caller.
void listtables_free(ListTables *lt)
{
    if (!lt) return;
    for (ListClass &lc : lt->cls) { b200_big_free(lc.d_tprim); b200_big_free(lc.d_tgeom); b200_big_free(lc.d_tnpp); }
    b200_dfree(lt->d_buf); b200_dfree(lt->d_cls_of); b200_dfree(lt->d_row_of); b200_dfree(lt->d_sdim); b200_dfree(lt->d_per); b200_big_free(lt->d_work);
    delete lt;
}

int listtables_build(CINTOpt *c)
{
    ListTables *lt = new ListTables();
    const size_t np = c->pairs.size();
    lt->cls_of.assign(np, -1); lt->row_of.assign(np, -1);
    std::map<std::vector<int>, int> key2class;
    for (size_t p = 0; p < np; p++) {
        const PairHdr &h = c->pairs[p];
        std::vector<int> key = {h.la, h.lb, h.nca, h.ncb};
        auto it = key2class.find(key);
        int ci;
        if (it == key2class.end()) {
            ci = (int)lt->cls.size();
            key2class[key] = ci;
            ListClass lc;
            lc.la = h.la; lc.lb = h.lb; lc.nca = h.nca; lc.ncb = h.ncb; lc.Q = 1;
            lt->cls.push_back(lc);
        } else ci = it->second;
        lt->cls_of[p] = ci;
        lt->row_of[p] = (int)lt->cls[ci].ids.size();
        lt->cls[ci].ids.push_back((int)p);
        lt->cls[ci].Q = std::max(lt->cls[ci].Q, h.npp);
    }
    const int ncls = (int)lt->cls.size();
    lt->choice.resize((size_t)ncls * ncls);
    lt->choice_cart.resize((size_t)ncls * ncls);
    for (int cart = 0; cart < 2; cart++)
    for (int cb = 0; cb < ncls; cb++)
        for (int ck = 0; ck < ncls; ck++) {
            const ListClass &B = lt->cls[cb], &K = lt->cls[ck];
            ListChoice &ch = (cart ? lt->choice_cart : lt->choice)[(size_t)cb * ncls + ck];
            memset(&ch, 0, sizeof ch);
            const int rs = c->omega != 0;
            ch.fn = reg_kernel_lookup(B.la, B.lb, K.la, K.lb, B.nca * B.ncb, K.nca * K.ncb, rs, cart);
            if (!ch.fn) { ch.fn = coop_kernel_lookup(B.la, B.lb, K.la, K.lb, B.nca * B.ncb, K.nca * K.ncb, &ch.ci, rs, cart); ch.coop = ch.fn != nullptr; }
            const int nroots_ = (B.la + B.lb + K.la + K.lb) / 2 + 1, ncu_ = K.nca * K.ncb, umax_ = std::max(1, K.Q);
            if (ch.fn && (ch.coop ? coop_kernel_smem(ch.ci, ncu_, umax_) : reg_kernel_smem(ch.fn, nroots_, ncu_, umax_)) > (size_t)tile_smem_limit()) {
                ch.fn = nullptr; ch.coop = 0;       // ket too deeply contracted for the staged primitives: generic kernel
            }
        }
    c->ltab = lt;
    return 0;
}

// structure-of-arrays primitive table of one bra class, uploaded the first time a list uses the class
int listclass_upload(CINTOpt *c, ListClass &lc)
{
    if (lc.d_tprim) return 0;
    {
        const size_t NT = lc.ids.size();
        const int nct = lc.nca * lc.ncb, Q = lc.Q;
        std::vector<double> tprim((size_t)(6 + nct) * Q * NT), tgeom(6 * NT);
        std::vector<int> nppc(NT);
        for (size_t n = 0; n < NT; n++) {
            const PairHdr &h = c->pairs[lc.ids[n]];
            fill_tprim(c, h, n, NT, Q, nct, tprim, tgeom);
            nppc[n] = std::max(h.npp, 1);
        }
        // pooled allocations (engine.cu:b200_big_alloc): a context per geometry step must not pay ~100 cudaMalloc / cudaFree pairs
        void *p0 = nullptr, *p1 = nullptr, *p2 = nullptr;
        if (b200_big_alloc(&p0, sizeof(double) * std::max<size_t>(1, tprim.size())) || b200_big_alloc(&p1, sizeof(double) * std::max<size_t>(1, tgeom.size())) ||
            b200_big_alloc(&p2, sizeof(int) * std::max<size_t>(1, nppc.size()))) return b200_fail(CINTB200_ENOMEM, "list-mode class tables");
        lc.d_tprim = (double *)p0; lc.d_tgeom = (double *)p1; lc.d_tnpp = (int *)p2;
        if (cudaMemcpy(lc.d_tprim, tprim.data(), sizeof(double) * tprim.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(lc.d_tgeom, tgeom.data(), sizeof(double) * tgeom.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(lc.d_tnpp, nppc.data(), sizeof(int) * nppc.size(), cudaMemcpyHostToDevice) != cudaSuccess)
            return b200_fail(CINTB200_ENODEV, "list-mode class tables: upload failed");
    }
    return 0;
}

// Evaluate the tuples whose (bra class, ket class) has a specialised kernel; handled[t] = 1 for those.  tasks are in the
// caller's order (Task::off = element offset of the tuple's block in d_out, strides for the packed block).  Spherical,
// plain Coulomb only.  Caller holds c->mtx; launches go to c->stream (not synchronised here).
int list_mode_run(CINTOpt *c, const Task *tasks, size_t n, double *d_out, unsigned char *handled, int cart)
{
    if (!c->ltab && listtables_build(c)) return CINTB200_ENOMEM;
    ListTables *lt = c->ltab;
    typedef ListChoice Choice;
    const int ncls = (int)lt->cls.size();
    if ((unsigned long long)ncls * ncls >= (1ull << 22) || n >= (1ull << 32)) return b200_fail(CINTB200_EINVAL, "list too large for the sort key");
    std::vector<int> gkey(n);                       // bra class * ncls + ket class
    // sort key: group | ket pair | orientation of the ket block | 127 - primitive count of the bra  (one 64-bit compare)
    struct SortRec { unsigned long long key; unsigned int t; };
    std::vector<SortRec> all(n);
#pragma omp parallel for schedule(static) if (n > 20000)
    for (long long t = 0; t < (long long)n; t++) {
        const int g = lt->cls_of[tasks[t].bra] * ncls + lt->cls_of[tasks[t].ket];
        gkey[t] = g;
        all[t].t = (unsigned int)t;
        all[t].key = ~0ull;
        if ((cart ? lt->choice_cart : lt->choice)[g].fn) {
            handled[t] = 1;
            const unsigned long long np_ = (unsigned long long)(127 - std::min(c->pairs[tasks[t].bra].npp, 127));
            all[t].key = ((unsigned long long)g << 42) | ((unsigned long long)(unsigned int)tasks[t].ket << 8)
                         | ((unsigned long long)(tasks[t].sc > tasks[t].sd) << 7) | np_;
        }
    }
    std::vector<SortRec> recs;
    recs.reserve(n);
    for (size_t t = 0; t < n; t++) if (handled[t]) recs.push_back(all[t]);
    all.clear(); all.shrink_to_fit();
    if (recs.empty()) return 0;
    auto rec_less = [](const SortRec &x, const SortRec &y) { return x.key != y.key ? x.key < y.key : x.t < y.t; };
    if (recs.size() > 50000) __gnu_parallel::sort(recs.begin(), recs.end(), rec_less);
    else std::sort(recs.begin(), recs.end(), rec_less);
    const size_t m = recs.size();
    std::vector<size_t> idx(m);
    for (size_t q = 0; q < m; q++) idx[q] = recs[q].t;
    auto choice = [&](int g) -> const Choice & { return (cart ? lt->choice_cart : lt->choice)[g]; };
    std::vector<int> tsel(m), tstride(2 * m), upair, ustride;
    std::vector<long long> trow(m);
    std::vector<int4> items;
    struct Group { size_t t0, t1, u0, u1, i0, i1; int key; };
    std::vector<Group> groups;
    size_t g0 = 0;
    while (g0 < m) {
        size_t g1 = g0;
        while (g1 < m && gkey[idx[g1]] == gkey[idx[g0]]) g1++;
        Group G;
        G.t0 = g0; G.t1 = g1; G.key = gkey[idx[g0]]; G.u0 = upair.size(); G.i0 = items.size();
        const Choice &ch = choice(G.key);
        const int per = ch.coop ? 32 / ch.ci.fs : 32;
        const size_t ng = g1 - g0;
        size_t k0 = g0;
        while (k0 < g1) {
            size_t k1 = k0;
            const Task &a = tasks[idx[k0]];
            while (k1 < g1 && tasks[idx[k1]].ket == a.ket && tasks[idx[k1]].sc == a.sc && tasks[idx[k1]].sd == a.sd) k1++;
            const int u = (int)(upair.size() - G.u0);
            upair.push_back(a.ket);
            ustride.push_back((int)a.sc); ustride.push_back((int)a.sd);        // re-laid out per group below
            for (size_t q = k0; q < k1; q += per)
                items.push_back(make_int4(u, (int)(q - g0), (int)std::min<size_t>(per, k1 - q), 0));
            k0 = k1;
        }
#pragma omp parallel for schedule(static) if (g1 - g0 > 20000)
        for (long long q = (long long)g0; q < (long long)g1; q++) {
            const Task &a = tasks[idx[q]];
            tsel[q] = lt->row_of[a.bra];
            trow[q] = a.off;
            tstride[2 * g0 + (q - g0)] = a.sa;            // group-local layout [2][ng]
            tstride[2 * g0 + ng + (q - g0)] = a.sb;
        }
        G.u1 = upair.size(); G.i1 = items.size();
        groups.push_back(G);
        g0 = g1;
    }
    // ustride per group as [2][NUg]
    std::vector<int> ustr2(ustride.size());
    size_t maxu = 1;
    for (const Group &G : groups) {
        const size_t nu = G.u1 - G.u0;
        maxu = std::max(maxu, nu);
        for (size_t u = 0; u < nu; u++) { ustr2[2 * G.u0 + u] = ustride[2 * (G.u0 + u)]; ustr2[2 * G.u0 + nu + u] = ustride[2 * (G.u0 + u) + 1]; }
    }
    // one device buffer: trow | ucol zeros | items | tsel | tstride | upair | ustride | counters
    auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
    const size_t o_trow = 0, o_ucol = al(o_trow + sizeof(long long) * m), o_items = al(o_ucol + sizeof(long long) * maxu);
    const size_t o_tsel = al(o_items + sizeof(int4) * items.size()), o_tstr = al(o_tsel + sizeof(int) * m);
    const size_t o_upair = al(o_tstr + sizeof(int) * 2 * m), o_ustr = al(o_upair + sizeof(int) * upair.size());
    const size_t o_cnt = al(o_ustr + sizeof(int) * ustr2.size()), bytes = al(o_cnt + sizeof(unsigned int) * groups.size());
    if (lt->cap < bytes) {
        b200_dfree(lt->d_buf); lt->d_buf = nullptr; lt->cap = 0;
        if (b200_dmalloc(&lt->d_buf, bytes * 2) != cudaSuccess) return b200_fail(CINTB200_ENOMEM, "list-mode work arrays: %zu bytes", bytes * 2);
        lt->cap = bytes * 2;
    }
    char *d = (char *)lt->d_buf;
    cudaStream_t st = c->stream;
    CU_OK(cudaMemsetAsync(d + o_ucol, 0, sizeof(long long) * maxu, st));
    CU_OK(cudaMemsetAsync(d + o_cnt, 0, sizeof(unsigned int) * groups.size(), st));
    CU_OK(cudaMemcpyAsync(d + o_trow, trow.data(), sizeof(long long) * m, cudaMemcpyHostToDevice, st));
    CU_OK(cudaMemcpyAsync(d + o_items, items.data(), sizeof(int4) * items.size(), cudaMemcpyHostToDevice, st));
    CU_OK(cudaMemcpyAsync(d + o_tsel, tsel.data(), sizeof(int) * m, cudaMemcpyHostToDevice, st));
    CU_OK(cudaMemcpyAsync(d + o_tstr, tstride.data(), sizeof(int) * 2 * m, cudaMemcpyHostToDevice, st));
    CU_OK(cudaMemcpyAsync(d + o_upair, upair.data(), sizeof(int) * upair.size(), cudaMemcpyHostToDevice, st));
    CU_OK(cudaMemcpyAsync(d + o_ustr, ustr2.data(), sizeof(int) * ustr2.size(), cudaMemcpyHostToDevice, st));
    CU_OK(cudaStreamSynchronize(st));               // the host vectors go out of scope when we return
    for (size_t g = 0; g < groups.size(); g++) {
        const Group &G = groups[g];
        const Choice &ch = choice(G.key);
        if (listclass_upload(c, lt->cls[G.key / ncls])) return CINTB200_ENOMEM;
        const ListClass &B = lt->cls[G.key / ncls], &K = lt->cls[G.key % ncls];
        const int nroots = (B.la + B.lb + K.la + K.lb) / 2 + 1;
        TileParams P;
        memset(&P, 0, sizeof P);
        P.tprim = B.d_tprim; P.tgeom = B.d_tgeom; P.tnpp = B.d_tnpp;
        P.NT = (int)B.ids.size(); P.Q = B.Q;
        P.trow = (const long long *)(d + o_trow) + G.t0;
        P.tstride = (const int *)(d + o_tstr) + 2 * G.t0;
        P.NTs = (int)(G.t1 - G.t0);
        P.tsel = (const int *)(d + o_tsel) + G.t0;
        P.t_begin = 0; P.t_end = (int)(G.t1 - G.t0); P.nca_t = B.nca;
        P.upair = (const int *)(d + o_upair) + G.u0;
        P.ucol = (const long long *)(d + o_ucol);
        P.ustride = (const int *)(d + o_ustr) + 2 * G.u0;
        P.NU = P.NU_all = (int)(G.u1 - G.u0); P.u_step = 1; P.u_first = 0; P.nca_u = K.nca; P.umax = std::max(1, K.Q);
        P.out = d_out; P.row0 = 0; P.ld = 1;
        P.pairs = c->d_pairs; P.prims = c->d_prims; P.pcoef = c->d_pcoef;
        P.rys = (!ch.coop && REG_FAST_RYS && nroots <= RYS_FNMAX && nroots <= REG_FAST_NMAX) ? c->d_rys_fast + rys_fast_off(nroots)
                                                                                          : c->d_rys + rys_tab_off(nroots);
        P.rs_w2 = c->omega * c->omega; P.rs_sign = c->omega;      /* sign * |omega| */ P.rs_pass0 = c->omega > 0 ? 1 : 0;
        P.items = (const int4 *)(d + o_items) + G.i0;
        P.nitems = (long long)(G.i1 - G.i0);
        P.gx = 1;
        P.counter = (unsigned int *)(d + o_cnt) + g;
        P.batch = (int)std::max<long long>(1, std::min<long long>(16, P.nitems / (148 * 32 * 4)));
        if (ch.coop ? coop_kernel_launch(ch.fn, ch.ci, K.nca * K.ncb, P, 1, 1, st) : reg_kernel_launch(ch.fn, nroots, K.nca * K.ncb, P, 1, 1, st))
            return b200_fail(CINTB200_ENODEV, "list-mode kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        c->launches++;
    }
    return 0;
}

// ------------------------------------------------------------------ dense shell-slice blocks
// out[i + NI (j + NJ (k + NK l))] for ALL shells i in [i0,i1), j in [j0,j1), k in [k0,k1), l in [l0,l1) (AO indices relative to
// the slice starts, column-major): the shape in which the reference's callers consume integrals (pyscf's fill drivers
// loop over shell slices and call the per-quartet function for every combination).  Seen as a tile, rows = AO pairs
// (i,j) with leading dimension NI*NJ and columns = AO pairs (k,l): the bra / ket lists are explicit, every ket meets
// every bra, and the tile kernels run unchanged (a pair's block is addressed through its own offset and strides).
struct RectEntry { int pair; long long off; int s_a, s_b; };      // pair id, block offset (rows or columns), strides of a, b

static int build_rect_plan(CINTOpt *c, JobPlan *plan, const std::vector<RectEntry> &T, const std::vector<RectEntry> &U,
                           long long ld, long long ncols, double *dev_out)
{
    struct ArenaScope { ArenaScope(DeviceArena *a) { g_arena = a; } ~ArenaScope() { if (g_arena) g_arena->flush(); g_arena = nullptr; } } arena_scope(&plan->arena);
    auto group = [&](const std::vector<RectEntry> &E, std::vector<PairClass> &out, std::vector<std::vector<int>> &members) {
        std::map<std::vector<int>, int> key2class;
        for (size_t n = 0; n < E.size(); n++) {
            const PairHdr &h = c->pairs[E[n].pair];
            std::vector<int> key = {h.la, h.lb, h.nca, h.ncb};
            auto it = key2class.find(key);
            int ci;
            if (it == key2class.end()) {
                ci = (int)out.size();
                key2class[key] = ci;
                PairClass pc;
                pc.la = h.la; pc.lb = h.lb; pc.nca = h.nca; pc.ncb = h.ncb; pc.Q = 1;
                out.push_back(pc);
                members.push_back({});
            } else ci = it->second;
            members[ci].push_back((int)n);
            out[ci].Q = std::max(out[ci].Q, h.npp);
        }
    };
    std::vector<std::vector<int>> tm, um;
    group(T, plan->classes, tm);
    group(U, plan->uclasses, um);
    plan->chunks.push_back({0, 1});
    plan->rows_before = {0, ld};
    plan->chunk_cols.push_back(ncols);
    plan->out_doubles = (size_t)ld * (size_t)ncols;
    for (size_t ci = 0; ci < plan->classes.size(); ci++) {
        PairClass &pc = plan->classes[ci];
        std::vector<int> &mem = tm[ci];
        // descending primitive count: neighbouring threads do equal work, a warp's first quartet carries its loop bound
        std::stable_sort(mem.begin(), mem.end(), [&](int x, int y) { return c->pairs[T[x].pair].npp > c->pairs[T[y].pair].npp; });
        const size_t NT = mem.size();
        const int nct = pc.nca * pc.ncb, Q = pc.Q;
        std::vector<double> tprim((size_t)(6 + nct) * Q * NT), tgeom(6 * NT);
        std::vector<long long> trow(NT);
        std::vector<int> tstride(2 * NT), nppc(NT);
        for (size_t n = 0; n < NT; n++) {
            const RectEntry &e = T[mem[n]];
            const PairHdr &h = c->pairs[e.pair];
            pc.ids.push_back(e.pair); pc.I.push_back(0); pc.npp.push_back(h.npp);
            fill_tprim(c, h, n, NT, Q, nct, tprim, tgeom);
            tstride[n] = e.s_a; tstride[NT + n] = e.s_b;
            trow[n] = e.off;
            nppc[n] = std::max(h.npp, 1);
        }
        pc.npp_prefix.assign(NT + 1, 0);
        for (size_t n = 0; n < NT; n++) pc.npp_prefix[n + 1] = pc.npp_prefix[n] + pc.npp[n];
        pc.chunk_lo = {0, (int)NT};
        if (upload(&pc.d_tprim, tprim) || upload(&pc.d_tgeom, tgeom) || upload(&pc.d_trow, trow) || upload(&pc.d_tstride, tstride) ||
            upload(&pc.d_tI, pc.I) || upload(&pc.d_tpair, pc.ids) || upload(&pc.d_tnpp, nppc))
            return CINTB200_ENOMEM;
    }
    for (size_t ci = 0; ci < plan->uclasses.size(); ci++) {
        PairClass &pc = plan->uclasses[ci];
        const std::vector<int> &mem = um[ci];
        const size_t NU = mem.size();
        std::vector<long long> ucol(NU);
        std::vector<int> ustride(2 * NU);
        for (size_t n = 0; n < NU; n++) {
            const RectEntry &e = U[mem[n]];
            pc.ids.push_back(e.pair); pc.I.push_back(0); pc.npp.push_back(c->pairs[e.pair].npp);
            ucol[n] = e.off; ustride[n] = e.s_a; ustride[NU + n] = e.s_b;
        }
        pc.npp_prefix.assign(NU + 1, 0);
        for (size_t n = 0; n < NU; n++) pc.npp_prefix[n + 1] = pc.npp_prefix[n] + pc.npp[n];
        pc.chunk_lo = {(int)NU, (int)NU};
        if (upload(&pc.d_tpair, pc.ids) || upload(&pc.d_tI, pc.I) || upload(&pc.d_ucol, ucol) || upload(&pc.d_ustride, ustride))
            return CINTB200_ENOMEM;
    }
    if (dev_out) { plan->d_out[0] = dev_out; plan->own_out = 0; }
    else if (b200_big_alloc((void **)&plan->d_out[0], sizeof(double) * std::max<size_t>(1, plan->out_doubles)))
        return b200_fail(CINTB200_ENOMEM, "cannot allocate the %zu-byte block buffer", sizeof(double) * plan->out_doubles);
    CU_OK(cudaStreamCreateWithFlags(&plan->copy_stream, cudaStreamNonBlocking));
    for (int b = 0; b < 2; b++) {
        CU_OK(cudaEventCreateWithFlags(&plan->ev_done[b], cudaEventDisableTiming));
        CU_OK(cudaEventCreateWithFlags(&plan->ev_copied[b], cudaEventDisableTiming));
    }
    for (int k = 0; k < JobPlan::NS; k++) {
        CU_OK(cudaStreamCreateWithFlags(&plan->streams[k], cudaStreamNonBlocking));
        CU_OK(cudaEventCreateWithFlags(&plan->ev_join[k], cudaEventDisableTiming));
    }
    CU_OK(cudaEventCreateWithFlags(&plan->ev_fork, cudaEventDisableTiming));
    CU_OK(cudaEventCreate(&plan->ev_t0));
    CU_OK(cudaEventCreate(&plan->ev_t1));
    return 0;
}

// ncenter 4: shls_slice = {i0,i1, j0,j1, k0,k1, l0,l1};  ncenter 3: {i0,i1, j0,j1, k0,k1}
int run_block(cintb200_ctx *c, int ncenter, const int *sl, double *out, int on_device, double *stats, int cart)
{
    if (!c || c->magic != B200_CTX_MAGIC) return b200_fail(CINTB200_EINVAL, "invalid context");
    if (!sl || !out) return b200_fail(CINTB200_EINVAL, "NULL shls_slice/out");
    for (int m = 0; m < ncenter; m++)
        if (sl[2 * m] < 0 || sl[2 * m + 1] > c->nbas || sl[2 * m] >= sl[2 * m + 1])
            return b200_fail(CINTB200_EINVAL, "shell slice %d = [%d, %d) is empty or outside 0..%d", m, sl[2 * m], sl[2 * m + 1], c->nbas);
    std::lock_guard<std::mutex> lock(c->mtx);
    CU_OK(cudaSetDevice(c->device));
    auto ao0 = [&](int sh) { return (long long)(cart ? c->shells[sh].ao_cart : c->shells[sh].ao_sph); };
    auto aoend = [&](int sh) { return ao0(sh) + (cart ? B200_NCART(c->shells[sh].l) : 2 * c->shells[sh].l + 1) * c->shells[sh].nctr; };
    const size_t npair2 = (size_t)c->nbas * (c->nbas + 1) / 2;
    std::vector<RectEntry> T, U;
    if (ncenter == 2) {
        // (i|k): bras and kets are the single-shell pseudo pairs; rows = AOs of the i slice, columns = AOs of the k slice
        const long long NI2 = aoend(sl[1] - 1) - ao0(sl[0]), NK2 = aoend(sl[3] - 1) - ao0(sl[2]);
        for (int i = sl[0]; i < sl[1]; i++) T.push_back(RectEntry{(int)(npair2 + i), ao0(i) - ao0(sl[0]), 1, 0});
        for (int k = sl[2]; k < sl[3]; k++) U.push_back(RectEntry{(int)(npair2 + k), ao0(k) - ao0(sl[2]), 1, 0});
        if (c->plan) { cudaDeviceSynchronize(); jobplan_free(c->plan); c->plan = nullptr; }
        JobPlan *plan2 = new JobPlan();
        plan2->ncenter = 3; plan2->rect = 2; plan2->force_generic = c->force_generic; plan2->schwarz_thr = 0; plan2->cart = cart;
        int rc2 = build_rect_plan(c, plan2, T, U, NI2, NK2, on_device ? out : nullptr);
        if (!rc2) rc2 = build_launches(c, plan2);
        if (rc2) { jobplan_free(plan2); return rc2; }
        c->plan = plan2;
        { TileSink ts; double *one[1] = {out}; if (!on_device) { ts.sinks = one; ts.nsinks = 1; }
          return execute_plan(c, plan2, 3, ts, stats); }      // 2- and 3-centre integrals share the cutoff
    }
    // a pair with virtual segmented copies (engine.cu:build_pairs) enters as nca x ncb single-contraction sub-blocks
    auto dim1 = [&](int sh) { return cart ? B200_NCART(c->shells[sh].l) : 2 * c->shells[sh].l + 1; };
    auto push_entry = [&](std::vector<RectEntry> &V, const RectEntry &e) {
        const int vf = (e.pair < (int)c->vfirst.size() && !c->force_generic) ? c->vfirst[e.pair] : -1;
        if (vf < 0) { V.push_back(e); return; }
        const PairHdr &h = c->pairs[e.pair];
        const long long da = dim1(h.sh_a), db = dim1(h.sh_b);
        for (int cb = 0; cb < h.ncb; cb++)
            for (int ca = 0; ca < h.nca; ca++) {
                RectEntry v = e;
                v.pair = vf + cb * h.nca + ca;
                v.off = e.off + ca * da * e.s_a + cb * db * e.s_b;
                V.push_back(v);
            }
    };
    const long long NI = aoend(sl[1] - 1) - ao0(sl[0]), NJ = aoend(sl[3] - 1) - ao0(sl[2]);
    const long long NK = aoend(sl[5] - 1) - ao0(sl[4]), NL = (ncenter == 4) ? aoend(sl[7] - 1) - ao0(sl[6]) : 1;
    if (NI * NJ > 0x7fffffffLL) return b200_fail(CINTB200_EINVAL, "bra slice too large: NI*NJ = %lld rows exceed 2^31", NI * NJ);
    for (int j = sl[2]; j < sl[3]; j++)
        for (int i = sl[0]; i < sl[1]; i++) {
            RectEntry e;
            e.pair = (int)((i >= j) ? (size_t)i * (i + 1) / 2 + j : (size_t)j * (j + 1) / 2 + i);
            e.off = (ao0(i) - ao0(sl[0])) + NI * (ao0(j) - ao0(sl[2]));
            const bool a_is_i = (c->pairs[e.pair].sh_a == i);         // i == j: a = b, the first index is a
            e.s_a = a_is_i ? 1 : (int)NI; e.s_b = a_is_i ? (int)NI : 1;
            push_entry(T, e);
        }
    if (ncenter == 4) {
        for (int l = sl[6]; l < sl[7]; l++)
            for (int k = sl[4]; k < sl[5]; k++) {
                RectEntry e;
                e.pair = (int)((k >= l) ? (size_t)k * (k + 1) / 2 + l : (size_t)l * (l + 1) / 2 + k);
                e.off = (ao0(k) - ao0(sl[4])) + NK * (ao0(l) - ao0(sl[6]));
                const bool a_is_k = (c->pairs[e.pair].sh_a == k);
                e.s_a = a_is_k ? 1 : (int)NK; e.s_b = a_is_k ? (int)NK : 1;
                push_entry(U, e);
            }
    } else {
        for (int k = sl[4]; k < sl[5]; k++) {
            RectEntry e;
            e.pair = (int)(npair2 + k);
            e.off = ao0(k) - ao0(sl[4]);
            e.s_a = 1; e.s_b = 0;
            U.push_back(e);
        }
    }
    double tph = b200_now();
    if (c->plan) { cudaDeviceSynchronize(); jobplan_free(c->plan); c->plan = nullptr; }
    b200_phase("block: release previous plan", tph);
    tph = b200_now();
    JobPlan *plan = new JobPlan();
    plan->ncenter = 3;                  // rectangular job: separate ket classes, every ket meets every bra (see build_launches)
    plan->rect = ncenter; plan->aux0 = 0; plan->rank = 0; plan->nranks = 1; plan->chunk_bytes = 0;
    plan->force_generic = c->force_generic; plan->schwarz_thr = 0; plan->cart = cart;
    int rc = build_rect_plan(c, plan, T, U, NI * NJ, NK * NL, on_device ? out : nullptr);
    b200_phase("block: build_rect_plan", tph);
    tph = b200_now();
    if (!rc) rc = build_launches(c, plan);
    b200_phase("block: build_launches", tph);
    if (rc) { jobplan_free(plan); return rc; }
    c->plan = plan;
    tph = b200_now();
    { TileSink ts; double *one[1] = {out}; if (!on_device) { ts.sinks = one; ts.nsinks = 1; }
      rc = execute_plan(c, plan, ncenter, ts, stats); }
    b200_phase("block: execute", tph);
    return rc;
}

extern "C" int cintb200_int2e_sph_block(cintb200_ctx *c, const int *shls_slice, double *out, int on_device, double *stats)
{ return run_block(c, 4, shls_slice, out, on_device, stats); }
extern "C" int cintb200_int3c2e_sph_block(cintb200_ctx *c, const int *shls_slice, double *out, int on_device, double *stats)
{ return run_block(c, 3, shls_slice, out, on_device, stats); }
extern "C" int cintb200_int2c2e_sph_block(cintb200_ctx *c, const int *shls_slice, double *out, int on_device, double *stats)
{ return run_block(c, 2, shls_slice, out, on_device, stats); }
// Cartesian output (int2e_cart / int3c2e_cart / int2c2e_cart): the same kernels with the cart->sph stages compiled out
extern "C" int cintb200_int2e_cart_block(cintb200_ctx *c, const int *shls_slice, double *out, int on_device, double *stats)
{ return run_block(c, 4, shls_slice, out, on_device, stats, 1); }
extern "C" int cintb200_int3c2e_cart_block(cintb200_ctx *c, const int *shls_slice, double *out, int on_device, double *stats)
{ return run_block(c, 3, shls_slice, out, on_device, stats, 1); }
extern "C" int cintb200_int2c2e_cart_block(cintb200_ctx *c, const int *shls_slice, double *out, int on_device, double *stats)
{ return run_block(c, 2, shls_slice, out, on_device, stats, 1); }

// Copy a rectangle of the most recent tile of chunk `chunk` ... (verification helper for tests):
// evaluates ONE chunk and returns it on the host together with its geometry.
extern "C" int cintb200_debug_chunk(cintb200_ctx *c, int chunk, double *host_out, size_t host_cap, long long *geom)
{
    if (!c || c->magic != B200_CTX_MAGIC || !c->plan) return b200_fail(CINTB200_EINVAL, "run cintb200_int2e_sph_all_unique first");
    JobPlan *plan = c->plan;
    if (chunk < 0 || chunk >= (int)plan->chunks.size()) return b200_fail(CINTB200_EINVAL, "chunk out of range");
    const int i0 = plan->chunks[chunk].first, i1 = plan->chunks[chunk].second;
    geom[0] = i0; geom[1] = i1;
    geom[2] = plan->rows_before[i0]; geom[3] = plan->rows_before[i1] - plan->rows_before[i0];
    geom[4] = plan->chunk_cols[chunk];
    geom[5] = (long long)plan->chunks.size();
    const size_t n = (size_t)geom[3] * (size_t)geom[4];
    if (!host_out) return 0;
    if (n > host_cap) return b200_fail(CINTB200_EINVAL, "host buffer too small");
    // chunks alternate between the two buffers in evaluation order
    // without a host sink every chunk is evaluated into buffer 0, so only the LAST chunk is still resident
    const int buf = 0;
    if (chunk != (int)plan->chunks.size() - 1) return b200_fail(CINTB200_EINVAL, "only the last chunk stays resident");
    CU_OK(cudaSetDevice(c->device));
    CU_OK(cudaMemcpy(host_out, plan->d_out[buf], sizeof(double) * n, cudaMemcpyDeviceToHost));
    return 0;
}

// column offset (this rank's numbering) and row offset of a pair, for tests
extern "C" int cintb200_debug_pair_offsets(cintb200_ctx *c, int i, int j, long long *row, long long *col_owner_rank)
{
    if (!c || c->magic != B200_CTX_MAGIC || !c->plan) return b200_fail(CINTB200_EINVAL, "no plan");
    JobPlan *plan = c->plan;
    if (i < j) std::swap(i, j);
    const int p = i * (i + 1) / 2 + j;
    if ((size_t)p >= plan->rowoff.size()) return b200_fail(CINTB200_EINVAL, "pair (%d,%d) is not a bra pair of the cached job", i, j);
    *row = plan->rowoff[p];
    *col_owner_rank = plan->colof[p];       // -1: another rank owns this ket
    return 0;
}

// 3-centre jobs: this rank's column offset of auxiliary shell k (-1: another rank owns it)
extern "C" int cintb200_debug_aux_offset(cintb200_ctx *c, int k, long long *col)
{
    if (!c || c->magic != B200_CTX_MAGIC || !c->plan || c->plan->ncenter != 3) return b200_fail(CINTB200_EINVAL, "no 3-centre plan");
    JobPlan *plan = c->plan;
    if (k < plan->aux0 || k >= c->nbas) return b200_fail(CINTB200_EINVAL, "shell %d is not an auxiliary shell", k);
    *col = plan->colof_aux[k - plan->aux0];
    return 0;
}

// Per-class profile of the last cintb200_int2e_sph_all_unique run made with profiling on: rows of
// 12 doubles {la, lb, lc, ld, nct, ncu, is_register_kernel, ms, quartets, primitive quartets, model flops, launches}.
extern "C" void cintb200_debug_profile(cintb200_ctx *c, int on) { if (c && c->magic == B200_CTX_MAGIC) c->profile = on; }
extern "C" int cintb200_debug_profile_rows(cintb200_ctx *c, double *rows, int max_rows)
{
    if (!c || c->magic != B200_CTX_MAGIC) return -1;
    const int n = (int)(c->profile_rows.size() / 12);
    if (rows) memcpy(rows, c->profile_rows.data(), sizeof(double) * 12 * std::min(n, max_rows));
    return n;
}

// Host-only planning: the static sharding of the whole job for `rank` of `nranks`, no GPU needed.
//   out[0] shell quartets   out[1] integrals   out[2] primitive quartets   out[3] model flops
//   out[4] columns owned    out[5] total rows  out[6] chunks               out[7] kernel launches
//   out[8] tile buffer bytes
static int plan_summary(int ncenter, int aux0, const int *atm, int natm, const int *bas, int nbas, const double *env,
                        int rank, int nranks, size_t chunk_bytes, double *out)
{
    if (nranks < 1 || rank < 0 || rank >= nranks || !out) return b200_fail(CINTB200_EINVAL, "bad rank %d of %d", rank, nranks);
    if (ncenter == 3 && (aux0 < 1 || aux0 >= nbas)) return b200_fail(CINTB200_EINVAL, "first auxiliary shell %d outside 1..%d", aux0, nbas - 1);
    CINTOpt *c = nullptr;
    int rc = ctx_new_host(&c, atm, natm, bas, nbas, env);
    if (rc) return rc;
    JobPlan *plan = new JobPlan();
    plan->ncenter = ncenter; plan->aux0 = (ncenter == 3) ? aux0 : 0; nbas = (ncenter == 3) ? aux0 : nbas;
    plan->rank = rank; plan->nranks = nranks; plan->host_only = 1;
    plan->chunk_bytes = chunk_bytes ? chunk_bytes : (size_t)16 << 30;
    rc = build_plan(c, plan);
    if (!rc) rc = build_launches(c, plan);
    if (!rc) {
        out[0] = plan->st_quartets; out[1] = plan->st_integrals; out[2] = plan->st_prim; out[3] = plan->st_flops;
        out[4] = (double)plan->cols_before[nbas]; out[5] = (double)plan->rows_before[nbas];
        out[6] = (double)plan->chunks.size(); out[7] = (double)plan->launches.size();
        out[8] = (double)plan->out_doubles * 8;
    }
    jobplan_free(plan);
    cintb200_destroy(c);
    return rc;
}

extern "C" int cintb200_plan_summary(const int *atm, int natm, const int *bas, int nbas, const double *env,
                                     int rank, int nranks, size_t chunk_bytes, double *out)
{ return plan_summary(4, 0, atm, natm, bas, nbas, env, rank, nranks, chunk_bytes, out); }

extern "C" int cintb200_plan_summary_3c(const int *atm, int natm, const int *bas, int nbas, const double *env, int aux_shell0,
                                        int rank, int nranks, size_t chunk_bytes, double *out)
{ return plan_summary(3, aux_shell0, atm, natm, bas, nbas, env, rank, nranks, chunk_bytes, out); }

// Launch list of the cached plan in execution order: rows of 12 doubles
// {chunk, la, lb, lc, ld, nct, ncu, kind (0 generic, 1 register, 2 cooperative), part, quartets, integrals, model flops}.
extern "C" int cintb200_debug_launch_rows(cintb200_ctx *c, double *rows, int max_rows)
{
    if (!c || c->magic != B200_CTX_MAGIC || !c->plan) return -1;
    const int n = (int)c->plan->launches.size();
    for (int k = 0; k < n && k < max_rows && rows; k++) {
        const LaunchRec &L = c->plan->launches[k];
        double *r = rows + 12 * k;
        r[0] = L.chunk;
        for (int j = 0; j < 6; j++) r[1 + j] = L.key[j];
        r[7] = L.fn ? (L.coop ? 2 : 1) : 0; r[8] = L.part; r[9] = L.quartets; r[10] = L.integrals; r[11] = L.flops;
    }
    return n;
}

// Fetch an nrow x ncol rectangle (column-major, leading dimension nrow) of the LAST chunk's tile, still resident in
// device buffer 0 after a run without host sink: verification of full-size jobs without copying whole tiles.
extern "C" int cintb200_debug_block(cintb200_ctx *c, long long row, long long col, int nrow, int ncol, double *host_out)
{
    if (!c || c->magic != B200_CTX_MAGIC || !c->plan) return b200_fail(CINTB200_EINVAL, "run cintb200_int2e_sph_all_unique first");
    JobPlan *plan = c->plan;
    const int ch = (int)plan->chunks.size() - 1;
    const long long row0 = plan->rows_before[plan->chunks[ch].first];
    const long long ld = plan->rows_before[plan->chunks[ch].second] - row0;
    if (row < row0 || row + nrow > row0 + ld || col < 0 || col + ncol > plan->chunk_cols[ch])
        return b200_fail(CINTB200_EINVAL, "rectangle outside the last chunk");
    CU_OK(cudaSetDevice(c->device));
    CU_OK(cudaMemcpy2D(host_out, sizeof(double) * nrow, plan->d_out[0] + (row - row0) + col * ld, sizeof(double) * ld,
                       sizeof(double) * nrow, ncol, cudaMemcpyDeviceToHost));
    return 0;
}
