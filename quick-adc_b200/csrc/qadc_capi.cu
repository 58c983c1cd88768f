// qadc_capi.cu — context, orchestration and the C ABI of libqadc_b200.so (include/qadc_b200.h).
// Host-side mirror of what scanner_4 (db_query_4.cpp:73-310) and the engines
// (query_common.hpp:149-309) do around the kernels; no CPU compute path exists here.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/qadc_b200.h"
#include "qadc_scan.cuh"
#include "qadc_tables.cuh"
#include "qadc_adc.cuh"
#include "qadc_flatprep.cuh"

using namespace qadc;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

constexpr int kMaxSmem = 227 * 1024;
constexpr int kNW = 8;
constexpr int kMaxBatch = 32768;   // queries per internal sub-batch of a search call
#ifndef QADC_NW1
#define QADC_NW1 15  // consumer warps of the single-query 16x4 flat kernel
#endif
#ifndef QADC_NS1
#define QADC_NS1 4   // ring stages of the single-query 16x4 flat kernel
#endif

}  // namespace

struct qadc_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    // quantisers
    int dim = 0, m = 0, bits = 4;
    float* d_codebooks = nullptr;
    float* d_rotation = nullptr;
    int K = 0;
    float* d_centroids = nullptr;
    // database
    int parts = 0;
    bool has_labels = false, begun = false, finalized = false;
    std::vector<uint32_t> h_size, h_pos_base, h_start_size, h_explicit_count;
    std::vector<uint64_t> h_sb_off, h_label_off, h_start_off;
    std::vector<uint8_t*> h_explicit_prefix;   // device pointers (owned, or into d_prefix_block) or null
    uint8_t* d_prefix_block = nullptr;         // qadc_set_prefixes: all explicit prefixes in one allocation
    uint64_t total_sb = 0, total_vec = 0;
    uint8_t* d_codes = nullptr;
    uint32_t* d_labels = nullptr;
    uint64_t *d_sb_off = nullptr, *d_label_off = nullptr, *d_start_off = nullptr;
    uint32_t *d_size = nullptr, *d_pos_base = nullptr, *d_start_size = nullptr;
    uint32_t* d_owned = nullptr;   // [parts] 1 = this shard answers for the partition ("owner computes"); default: size > 0
    uint8_t* d_starts = nullptr;
    uint8_t* d_starts_native = nullptr;   // flat database, long prefix: the prefix as nibble-plane superblocks (qadc_flatprep.cuh)
    uint32_t max_start = 0;
    // plain-ADC database (db_query path): row-major codes, independent of the Quick ADC layout above
    int adc_parts = 0;
    uint8_t* d_rows = nullptr;
    uint64_t* d_row_off = nullptr;    // [adc_parts + 1]
    uint32_t* d_adc_labels = nullptr;
    DevBuf b_adc_dists;
    // scratch
    DevBuf staging, b_queries, b_assign, b_tables, b_tmin, b_qmax, b_qmin, b_qtables, b_lists, b_plists, b_ids,
        b_dists, b_counts, b_keys, b_dump, b_hist, b_sbound, b_cdist, b_qmin_raw, b_prov, b_cand, b_ghist;
    int local_nq = 0, local_ma = 0, local_r = 0;   // batch whose tables qadc_tables_local_device left in the scratch
    bool local_mode = false;                        // scan_device runs for qadc_search_bounded_device: only owned probes have tables
    int* d_err = nullptr;
    int* h_err = nullptr;   // pinned
    // options / accounting
    long opt_flat_qb = 0, opt_flat_chunks = 0, opt_flat_filter = 1, opt_ivf_fused = 1, opt_flat_ring = 1, opt_flat_seed = 1, opt_flat_prep = 1, opt_seed_minus = 0, opt_flat_share = 1;
    bool sbound_seeded = false;   // the fused inverted-list table kernel already wrote the shared bounds of this batch
    bool ghist_zeroed = false;    // the table pipeline of this batch already zeroed b_ghist for the scan that follows
    int ivf_sb_per_item = 8;   // superblocks per work item of the IVF scan (option "ivf_sb_per_item")
    int launches = 0;
    cudaEvent_t ev[8] = {};
    static constexpr int kScanRing = 64;
    cudaEvent_t ev_scan0[kScanRing] = {}, ev_scan1[kScanRing] = {};   // ring: one pair per search call
    long scan_seq = 0;
    bool scan_timed = false;
};

namespace {

int fail(qadc_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    else g_create_error = msg;
    return code;
}

#define QCK(call)                                                                                      \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess)                                                                        \
            return fail(ctx, QADC_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__));         \
    } while (0)

int ensure(qadc_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return QADC_OK;
    if (b.p) QCK(cudaFree(b.p));
    b.p = nullptr; b.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    QCK(cudaMalloc(&b.p, want));
    b.cap = want;
    return QADC_OK;
}
#define ENSURE(buf, bytes)                                   \
    do {                                                     \
        int rc__ = ensure(ctx, buf, bytes);                  \
        if (rc__) return rc__;                               \
    } while (0)

PipeK make_pipek() {
    return PipeK{1u, 0xffffffffu, 1u, 1u << 8, 1u << 16, 1u << 24};
}

int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

void free_db(qadc_ctx* c) {
    cudaFree(c->d_codes); cudaFree(c->d_labels); cudaFree(c->d_sb_off); cudaFree(c->d_label_off);
    cudaFree(c->d_start_off); cudaFree(c->d_owned); cudaFree(c->d_size); cudaFree(c->d_pos_base); cudaFree(c->d_start_size);
    cudaFree(c->d_starts); cudaFree(c->d_starts_native); c->d_starts_native = nullptr;
    for (auto p : c->h_explicit_prefix)
        if (p && !c->d_prefix_block) cudaFree(p);
    cudaFree(c->d_prefix_block); c->d_prefix_block = nullptr;
    c->d_codes = nullptr; c->d_labels = nullptr; c->d_sb_off = c->d_label_off = c->d_start_off = nullptr;
    c->d_size = c->d_pos_base = c->d_start_size = c->d_owned = nullptr; c->d_starts = nullptr;
    c->h_explicit_prefix.clear(); c->h_explicit_count.clear();
    c->begun = c->finalized = false;
}

// ---- kernels that only the capi needs ------------------------------------------------------
// Row-major keep-prefix of every partition out of the native layout (scanner_4 keeps the same
// row-major copy in starts_flat, db_query_4.cpp:153-168).  grid = (K, y).
template <int M>
__global__ void extract_prefix_kernel(const uint8_t* __restrict__ native, const uint64_t* __restrict__ sb_off,
                                      const uint32_t* __restrict__ start_size, const uint64_t* __restrict__ start_off,
                                      const uint8_t* __restrict__ has_explicit, uint8_t* __restrict__ starts) {
    constexpr int CS = M / 2;
    const int p = blockIdx.x;
    if (has_explicit[p]) return;
    const uint32_t n = start_size[p];
    const uint32_t* base = reinterpret_cast<const uint32_t*>(native + sb_off[p] * sb_bytes(M));
    for (uint32_t v = blockIdx.y * blockDim.x + threadIdx.x; v < n; v += gridDim.y * blockDim.x) {
        const uint32_t sb = v / kSbVec, lane = (v % kSbVec) / 8, k = v % 8;
        const uint32_t* w = base + static_cast<size_t>(sb) * (sb_bytes(M) / 4);
        for (int b = 0; b < CS; ++b) {
            const int j0 = 2 * b, j1 = 2 * b + 1;
            const uint32_t lo = (w[((j0 / 4) * 32 + lane) * 4 + (j0 % 4)] >> (4 * k)) & 15u;
            const uint32_t hi = (w[((j1 / 4) * 32 + lane) * 4 + (j1 % 4)] >> (4 * k)) & 15u;
            starts[(start_off[p] + v) * CS + b] = static_cast<uint8_t>(lo | (hi << 4));
        }
    }
}

#ifdef QADC_EXPERIMENT
// tools/exp_seed.py (build with QADC_NVCC_EXTRA=-DQADC_EXPERIMENT): how much would a tighter shared-bound seed be worth
__global__ void seed_minus_kernel(int* sbound, int nq, int k) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) sbound[q] = max(0, sbound[q] - k);
}
#endif

// ---- flat scan dispatch ---------------------------------------------------------------------
template <int M, int QB, int NW, int NS>
int launch_flat(qadc_ctx* ctx, FlatScanArgs a, int chunks) {
    using Cfg = FlatCfg<M, QB, NW, NS>;
    const size_t smem = Cfg::smem_bytes(a.cap);
    if (smem > kMaxSmem) return QADC_ENOMEM;
    auto kern = scan_flat_kernel<M, QB, NW, NS>;
    QCK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    dim3 grid(chunks, (a.nq + QB - 1) / QB);
    kern<<<grid, Cfg::kThreads, smem, ctx->stream>>>(a);
    ctx->launches++;
    QCK(cudaGetLastError());
    return QADC_OK;
}

template <int NW, int NSW, int NPS>
int launch_flat_wr(qadc_ctx* ctx, FlatScanArgs a, int chunks) {
    using Cfg = WarpRingCfg<NW, NSW, NPS>;
    const size_t smem = Cfg::smem_bytes(a.cap);
    if (smem > kMaxSmem) return QADC_ENOMEM;
    auto kern = scan_flat_wr_kernel<NW, NSW, NPS>;
    QCK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    dim3 grid(chunks, a.nq);
    kern<<<grid, Cfg::kThreads, smem, ctx->stream>>>(a);
    ctx->launches++;
    QCK(cudaGetLastError());
    return QADC_OK;
}

template <int M, int QB, int NW, int NSW>
int launch_flat_wrq(qadc_ctx* ctx, FlatScanArgs a, int chunks) {
    using Cfg = WarpRingBatchCfg<M, QB, NW, NSW>;
    const size_t smem = Cfg::smem_bytes(a.cap);
    if (smem > kMaxSmem) return QADC_ENOMEM;
    auto kern = scan_flat_wrq_kernel<M, QB, NW, NSW>;
    QCK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    dim3 grid(chunks, (a.nq + QB - 1) / QB);
    kern<<<grid, Cfg::kThreads, smem, ctx->stream>>>(a);
    ctx->launches++;
    QCK(cudaGetLastError());
    return QADC_OK;
}

struct FlatVariant {
    int m, qb, nw, ns;
    int (*launch)(qadc_ctx*, FlatScanArgs, int);
    size_t (*smem)(int cap);
    int wr;   // > 0: per-warp rings (scan_flat_wr_kernel), chosen with option "flat_ring"; the value is the superblocks per ring slot
};
#define QADC_FLAT_VARIANT(M, QB, NW, NS) {M, QB, NW, NS, launch_flat<M, QB, NW, NS>, FlatCfg<M, QB, NW, NS>::smem_bytes, 0}
#ifndef QADC_NWR
#define QADC_NWR 16   // warps of the per-warp-ring kernel
#endif
#ifndef QADC_NSWR
#define QADC_NSWR 4   // ring slots per warp
#endif
#ifndef QADC_NPSR
#define QADC_NPSR 1   // superblocks per ring slot (one TMA copy and one barrier per slot)
#endif
// in order of preference per (m, qb): 15 consumer warps when the lists fit, else 8
const FlatVariant kFlatVariants[] = {
    {16, 1, QADC_NWR, QADC_NSWR * QADC_NPSR, launch_flat_wr<QADC_NWR, QADC_NSWR, QADC_NPSR>, WarpRingCfg<QADC_NWR, QADC_NSWR, QADC_NPSR>::smem_bytes, QADC_NPSR},
#define QADC_FLAT_WRQ(M, QB, NW, NSW) {M, QB, NW, NSW, launch_flat_wrq<M, QB, NW, NSW>, WarpRingBatchCfg<M, QB, NW, NSW>::smem_bytes, 1}
    QADC_FLAT_WRQ(16, 2, 16, 4), QADC_FLAT_WRQ(16, 4, 16, 2), QADC_FLAT_WRQ(32, 1, 16, 2), QADC_FLAT_WRQ(32, 2, 16, 2),
    QADC_FLAT_VARIANT(16, 1, QADC_NW1, QADC_NS1), QADC_FLAT_VARIANT(16, 1, 8, 4),
    QADC_FLAT_VARIANT(16, 2, 15, 3), QADC_FLAT_VARIANT(16, 2, 8, 4),
    QADC_FLAT_VARIANT(16, 4, 15, 3), QADC_FLAT_VARIANT(16, 4, 8, 4),
    QADC_FLAT_VARIANT(32, 1, 15, 3), QADC_FLAT_VARIANT(32, 1, 8, 4),
    QADC_FLAT_VARIANT(32, 2, 15, 2), QADC_FLAT_VARIANT(32, 2, 8, 4),
};

struct FlatPlan { const FlatVariant* v; int qb, nw, chunks, cap; uint32_t sb_per_chunk; };

// Chooses queries-per-pass, kernel variant, list capacity and chunk count for the flat scan.
int plan_flat(qadc_ctx* ctx, int nq, int r, FlatPlan& pl) {
    const int M = ctx->m;
    int qb = static_cast<int>(ctx->opt_flat_qb);
    // measured on config 1 (1M x 16x4, 10 000 queries, per-warp-ring batched kernel): 709 (4) / 669 (2) / 511 (1) G pairs/s
    if (qb <= 0) qb = (nq >= 4 && M == 16) ? 4 : (nq >= 2) ? 2 : 1;
    if (M == 32 && qb > 2) qb = 2;
    if (qb > 4) qb = 4;
    if (qb == 3) qb = 2;
    while (qb > nq && qb > 1) qb >>= 1;
    // list capacity: one superblock of candidates on top of r (one pass) or half a superblock (two passes)
    const int caps[2] = {next_pow2(r + kSbVec), next_pow2(r + kSbVec / 2)};
    pl.v = nullptr;
    for (; qb >= 1 && !pl.v; qb >>= 1)
        for (const FlatVariant& v : kFlatVariants) {
            if (v.m != M || v.qb != qb || pl.v || (v.wr && !ctx->opt_flat_ring)) continue;
            for (int cap : caps)
                if (!pl.v && v.smem(cap) <= static_cast<size_t>(kMaxSmem)) { pl.v = &v; pl.cap = cap; pl.qb = qb; pl.nw = v.nw; }
        }
    if (!pl.v) return fail(ctx, QADC_EINVAL, "r too large for the scan kernel's shared-memory lists");
    const uint32_t n_sb = static_cast<uint32_t>(ctx->total_sb);
    const int tile_sb = pl.nw * std::max(1, pl.v->wr);
    const int qgroups = (nq + pl.qb - 1) / pl.qb;
    long chunks = ctx->opt_flat_chunks;
    // up to ~16 waves of CTAs (the tail of the last wave then costs a few percent), fewer for small
    // shards so that a CTA still streams >= ~600 tiles and its ~20 us of setup/merge stays small
    if (chunks <= 0) {
        const long n_tiles = (n_sb + pl.nw - 1) / pl.nw;
        int waves = 16;
        while (waves > 1 && n_tiles / std::max<long>(1, (static_cast<long>(waves) * ctx->sm_count + qgroups - 1) / qgroups) < 600)
            waves >>= 1;
        chunks = std::max<long>(1, (static_cast<long>(waves) * ctx->sm_count + qgroups - 1) / qgroups);
    }
    const long max_chunks = std::max<long>(1, (n_sb + tile_sb - 1) / tile_sb);
    chunks = std::min(chunks, max_chunks);
    uint32_t spc = static_cast<uint32_t>((n_sb + chunks - 1) / chunks);
    spc = ((spc + tile_sb - 1) / tile_sb) * tile_sb;   // whole tiles per chunk
    pl.sb_per_chunk = std::max<uint32_t>(spc, tile_sb);
    pl.chunks = static_cast<int>((n_sb + pl.sb_per_chunk - 1) / pl.sb_per_chunk);
    if (pl.chunks < 1) pl.chunks = 1;
    return QADC_OK;
}

// grid.x of the int8 passes over a flat prefix: about eight CTAs (64 warps) per SM over all queries, so that a warp has
// only a few superblocks to fetch one after the other (at two CTAs per SM the pass was latency-bound: 22 us for the
// 500 000-vector prefix of the 1e9 bench and 16 queries)
unsigned flat_prep_splits(const qadc_ctx* ctx, uint32_t n_sb, int nq) {
    const unsigned want = static_cast<unsigned>((8 * ctx->sm_count + nq - 1) / nq);
    return std::max(1u, std::min(want, (n_sb + 7) / 8));
}

// Seeds the per-query shared bound from the keep-prefixes (prefix_hist + prefix_bound kernels).
int seed_shared_bound(qadc_ctx* ctx, const int32_t* d_assign, const int8_t* d_qtables, int nq, int ma, int r) {
    ENSURE(ctx->b_hist, static_cast<size_t>(nq) * 129 * 4);   // 128 bins per query + one ticket per query
    ENSURE(ctx->b_sbound, static_cast<size_t>(nq) * 4);
    QCK(cudaMemsetAsync(ctx->b_hist.p, 0, static_cast<size_t>(nq) * 129 * 4, ctx->stream));
    unsigned int* hist = ctx->b_hist.as<unsigned int>();
    if (ctx->K == 0 && ctx->d_starts_native && ctx->opt_flat_prep && ctx->opt_flat_seed) {
        // long flat prefix: the same histogram from the nibble-plane copy of the prefix with the PRMT core, its scan by the
        // last CTA of each query (qadc_flatprep.cuh): one launch
        const uint32_t n_prefix = ctx->h_start_size[0], n_sb = (n_prefix + kSbVec - 1) / kSbVec;
        dim3 pgrid(flat_prep_splits(ctx, n_sb, nq), nq);
        unsigned int* done = hist + static_cast<size_t>(nq) * 128;
        if (ctx->m == 16)
            flat_prefix_pass_kernel<16, 0><<<pgrid, 256, 0, ctx->stream>>>(ctx->d_starts_native, n_prefix, d_qtables, hist, done,
                                                                          ctx->b_sbound.as<int>(), 126, r, 0, nullptr, nullptr, 0u, make_pipek());
        else
            flat_prefix_pass_kernel<32, 0><<<pgrid, 256, 0, ctx->stream>>>(ctx->d_starts_native, n_prefix, d_qtables, hist, done,
                                                                          ctx->b_sbound.as<int>(), 126, r, 0, nullptr, nullptr, 0u, make_pipek());
        ctx->launches++;
        QCK(cudaGetLastError());
        return QADC_OK;
    }
    PrefixBoundArgs pa;
    pa.starts = ctx->d_starts; pa.start_off = ctx->d_start_off; pa.start_size = ctx->d_start_size;
    pa.assign = d_assign; pa.qtabs = d_qtables; pa.ma = ma;
    pa.nsplit = (ctx->K == 0) ? static_cast<int>(std::min<uint32_t>(64, std::max<uint32_t>(1, ctx->max_start / 8192))) : 1;
    pa.hist = hist;
    pa.owned_size = ctx->local_mode ? ctx->d_owned : nullptr;
    dim3 grid(pa.nsplit, nq);
    // flat_seed = 0 (flat databases): no histogram pass, the bound starts at 126 and the scan's own candidates tighten it
    if (ctx->K != 0 || ctx->opt_flat_seed) {
        if (ctx->m == 16) prefix_hist_kernel<16><<<grid, 256, 0, ctx->stream>>>(pa);
        else prefix_hist_kernel<32><<<grid, 256, 0, ctx->stream>>>(pa);
        ctx->launches++;
        QCK(cudaGetLastError());
    }
    prefix_bound_kernel<<<nq, 128, 0, ctx->stream>>>(pa.hist, r, ctx->b_sbound.as<int>());
    ctx->launches++;
    QCK(cudaGetLastError());
    return QADC_OK;
}

int run_merge(qadc_ctx* ctx, MergeArgs ma) {
    merge_lists_kernel<<<ma.nq, kMergeThreads, 0, ctx->stream>>>(ma);
    ctx->launches++;
    QCK(cudaGetLastError());
    return QADC_OK;
}

// Stage S on device buffers: assign [nq][ma], qtables [nq][ma][M*16] -> ids/dists/counts/keys.
int scan_device(qadc_ctx* ctx, const int32_t* d_assign, const int8_t* d_qtables, int nq, int ma, int r,
                uint32_t* d_ids, int8_t* d_dists, int32_t* d_counts, uint64_t* d_keys) {
    const int M = ctx->m;
    const bool flat = (ctx->K == 0);
    int n_lists = 0;
    const PipeK pk = make_pipek();
    int rc = QADC_OK;
    if (ctx->sbound_seeded) ctx->sbound_seeded = false;   // seeded by ivf_prepare_kernel for exactly this batch
    else rc = seed_shared_bound(ctx, d_assign, d_qtables, nq, ma, r);
    if (rc) return rc;
#ifdef QADC_EXPERIMENT
    if (ctx->opt_seed_minus) {
        seed_minus_kernel<<<(nq + 255) / 256, 256, 0, ctx->stream>>>(ctx->b_sbound.as<int>(), nq, static_cast<int>(ctx->opt_seed_minus));
        QCK(cudaGetLastError());
    }
#endif
    if (flat) {
        FlatPlan pl;
        rc = plan_flat(ctx, nq, r, pl);
        if (rc) return rc;
        // the NW warp lists of a CTA are merged inside the kernel when they fit a CTA-wide sort in the tile ring
        const size_t ring_keys = static_cast<size_t>(pl.v->ns) * pl.nw * sb_bytes(M) / 8;   // same product for both ring layouts
        const bool cta_merge = static_cast<size_t>(next_pow2(pl.nw * r)) <= std::min<size_t>(ring_keys, 4096);
        n_lists = cta_merge ? pl.chunks : pl.chunks * pl.nw;
        ENSURE(ctx->b_lists, static_cast<size_t>(nq) * n_lists * r * 8);
        FlatScanArgs a;
        a.codes = ctx->d_codes; a.n_sb = static_cast<uint32_t>(ctx->total_sb); a.size = ctx->h_size[0];
        a.pos_base = ctx->h_pos_base[0]; a.sb_per_chunk = pl.sb_per_chunk; a.qtabs = d_qtables; a.nq = nq;
        a.r = r; a.cap = pl.cap; a.lists = ctx->b_lists.as<uint64_t>(); a.n_lists = n_lists;
        a.shared_bound = ctx->b_sbound.as<int>(); a.k = pk; a.use_filter = static_cast<int>(ctx->opt_flat_filter);
        a.ghist = nullptr;
        if (ctx->opt_flat_share && pl.chunks > 1) {
            // the chunks of a query pool their candidates in a global histogram (hist_publish)
            if (!ctx->ghist_zeroed) {   // (the long-prefix table pipeline zeroes it together with its own counters)
                ENSURE(ctx->b_ghist, static_cast<size_t>(nq) * 129 * 4);
                QCK(cudaMemsetAsync(ctx->b_ghist.p, 0, static_cast<size_t>(nq) * 128 * 4, ctx->stream));
            }
            a.ghist = ctx->b_ghist.as<int>();
        }
        ctx->ghist_zeroed = false;
        if (ctx->scan_timed) QCK(cudaEventRecord(ctx->ev_scan0[ctx->scan_seq % qadc_ctx::kScanRing], ctx->stream));
        rc = pl.v->launch(ctx, a, pl.chunks);
        if (rc) return rc == QADC_ENOMEM ? fail(ctx, rc, "scan kernel shared memory exceeds 227 KB") : rc;
    } else {
        const int cap = next_pow2(r + kSbVec);
        int chunks = std::max(1, std::min((ma + kNW - 1) / kNW, (2 * ctx->sm_count + nq - 1) / nq));
        const int ppc = (ma + chunks - 1) / chunks;
        chunks = (ma + ppc - 1) / ppc;
        const size_t smem = ivf_smem_bytes(M, kNW, cap, ppc);
        if (smem > kMaxSmem) return fail(ctx, QADC_EINVAL, "r or ma too large for the IVF scan kernel");
        n_lists = chunks * kNW;
        ENSURE(ctx->b_lists, static_cast<size_t>(nq) * n_lists * r * 8);
        IvfScanArgs a;
        a.codes = ctx->d_codes; a.part_sb_off = ctx->d_sb_off; a.part_size = ctx->d_size;
        a.part_pos_base = ctx->d_pos_base; a.assign = d_assign; a.qtabs = d_qtables; a.nq = nq; a.ma = ma;
        a.r = r; a.cap = cap; a.probes_per_chunk = ppc; a.sb_per_item = ctx->ivf_sb_per_item;
        a.lists = ctx->b_lists.as<uint64_t>(); a.n_lists = n_lists;
        a.shared_bound = ctx->b_sbound.as<int>(); a.k = pk;
        dim3 grid(chunks, nq);
        if (ctx->scan_timed) QCK(cudaEventRecord(ctx->ev_scan0[ctx->scan_seq % qadc_ctx::kScanRing], ctx->stream));
        if (M == 16) {
            QCK(cudaFuncSetAttribute(scan_ivf_kernel<16, kNW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
            scan_ivf_kernel<16, kNW><<<grid, kNW * 32, smem, ctx->stream>>>(a);
        } else {
            QCK(cudaFuncSetAttribute(scan_ivf_kernel<32, kNW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
            scan_ivf_kernel<32, kNW><<<grid, kNW * 32, smem, ctx->stream>>>(a);
        }
        ctx->launches++;
        QCK(cudaGetLastError());
    }
    if (ctx->scan_timed) { QCK(cudaEventRecord(ctx->ev_scan1[ctx->scan_seq % qadc_ctx::kScanRing], ctx->stream)); ctx->scan_seq++; }
    MergeArgs mg{};
    mg.in_keys = ctx->b_lists.as<uint64_t>(); mg.in_ids = nullptr; mg.L = n_lists; mg.r = r; mg.nq = nq;
    mg.shard_major = 0; mg.out_keys = d_keys; mg.out_ids = d_ids; mg.out_dists = d_dists; mg.out_counts = d_counts;
    mg.out_rth_value = nullptr;
    mg.init_bound = ctx->b_sbound.as<int>();   // final shared bound: at least r scanned vectors are at or below it
    if (ctx->has_labels) {
        mg.labels = ctx->d_labels; mg.label_off = ctx->d_label_off; mg.part_pos_base = ctx->d_pos_base;
        mg.assign = d_assign; mg.ma = ma;
    }
    return run_merge(ctx, mg);
}

// Stage T on device buffers. d_assign_in may be null (then assignment is computed).
template <int M, int DSQ>
int launch_ivf_prepare(qadc_ctx* ctx, const IvfPrepArgs& a, int nq, size_t smem) {
    auto kern = ivf_prepare_kernel<M, DSQ>;
    QCK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<nq, 256, smem, ctx->stream>>>(a);
    ctx->launches++;
    QCK(cudaGetLastError());
    return QADC_OK;
}

int tables_device(qadc_ctx* ctx, const float* d_queries, int nq, int ma, int r, const int32_t* d_assign_in,
                  bool record_events, bool want_float_tables = false) {
    const int M = ctx->m, dim = ctx->dim;
    const bool flat = (ctx->K == 0);
    const size_t nqa = static_cast<size_t>(nq) * ma;
    // inverted lists: the whole pipeline after the assignment in one kernel, tables resident in shared memory
    const size_t fused_smem = ivf_prep_smem_bytes(M, ma, dim);
    // (two CTAs per SM or it loses to the separate kernels: measured 6.3 vs 5.7 ms on config 5, nprobe 128, where the
    // tables of one query take 128 KB; opt_ivf_fused = 2 forces it whenever it fits)
    const size_t fused_limit = ctx->opt_ivf_fused >= 2 ? static_cast<size_t>(kMaxSmem) - 2048 : 112 * 1024;
    const bool fused = !flat && ctx->opt_ivf_fused && fused_smem <= fused_limit &&
                       static_cast<uint64_t>(ma) * ctx->max_start <= (1u << 16);
    ENSURE(ctx->b_assign, nqa * 4);
    if (!fused || want_float_tables) ENSURE(ctx->b_tables, nqa * M * 16 * 4);
    ENSURE(ctx->b_tmin, nqa * 4);
    ENSURE(ctx->b_qmax, static_cast<size_t>(nq) * 4);
    ENSURE(ctx->b_qmin, static_cast<size_t>(nq) * 4);
    ENSURE(ctx->b_qtables, nqa * M * 16);
    int32_t* d_assign = ctx->b_assign.as<int32_t>();
    if (record_events) QCK(cudaEventRecord(ctx->ev[0], ctx->stream));
    // 1. assignment (flat_db: assign = 0, databases.hpp:93-101; index_db: :201-211)
    if (d_assign_in) {
        QCK(cudaMemcpyAsync(d_assign, d_assign_in, nqa * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    } else if (flat) {
        QCK(cudaMemsetAsync(d_assign, 0, nqa * 4, ctx->stream));
    } else {
        // distances for chunks of queries (<= 1 GiB of floats at a time), then the per-query selection
        const int chunk_q = std::max(1, static_cast<int>(std::min<size_t>(nq, (size_t(1) << 28) / ctx->K)));
        ENSURE(ctx->b_cdist, static_cast<size_t>(chunk_q) * ctx->K * 4);
        for (int q0 = 0; q0 < nq; q0 += chunk_q) {
            const int n = std::min(chunk_q, nq - q0);
            dim3 grid((n + kCoarseTQ - 1) / kCoarseTQ, (ctx->K + kCoarseTC - 1) / kCoarseTC);
            coarse_dist_kernel<<<grid, 256, 0, ctx->stream>>>(d_queries + static_cast<size_t>(q0) * dim, n, dim,
                                                              ctx->d_centroids, ctx->K, ctx->b_cdist.as<float>());
            QCK(cudaGetLastError());
            coarse_select_kernel<<<n, kSelThreads, 0, ctx->stream>>>(ctx->b_cdist.as<float>(), ctx->K, ma,
                                                                    d_assign + static_cast<size_t>(q0) * ma, nullptr, 0u);
            QCK(cudaGetLastError());
            ctx->launches += 2;
        }
    }
    if (record_events) QCK(cudaEventRecord(ctx->ev[1], ctx->stream));
    if (fused) {
        ENSURE(ctx->b_sbound, static_cast<size_t>(nq) * 4);
        IvfPrepArgs pa;
        pa.queries = d_queries; pa.dim = dim; pa.codebooks = ctx->d_codebooks; pa.rotation = ctx->d_rotation;
        pa.centroids = ctx->d_centroids; pa.assign = d_assign; pa.ma = ma; pa.r = r;
        pa.starts = ctx->d_starts; pa.start_off = ctx->d_start_off; pa.start_size = ctx->d_start_size;
        pa.tables_out = want_float_tables ? ctx->b_tables.as<float>() : nullptr;
        pa.qtables = ctx->b_qtables.as<int8_t>(); pa.qmin = ctx->b_qmin.as<float>(); pa.qmax = ctx->b_qmax.as<float>();
        pa.shared_bound = ctx->b_sbound.as<int>(); pa.err = ctx->d_err;
        int rc;
        const int dsq = dim / M;
#define QADC_PREP(MM, DD) rc = launch_ivf_prepare<MM, DD>(ctx, pa, nq, fused_smem)
        if (M == 16) {
            switch (dsq) {
                case 2: QADC_PREP(16, 2); break;
                case 4: QADC_PREP(16, 4); break;
                case 6: QADC_PREP(16, 6); break;
                case 8: QADC_PREP(16, 8); break;
                default: QADC_PREP(16, 0); break;
            }
        } else {
            switch (dsq) {
                case 2: QADC_PREP(32, 2); break;
                case 3: QADC_PREP(32, 3); break;
                case 4: QADC_PREP(32, 4); break;
                default: QADC_PREP(32, 0); break;
            }
        }
#undef QADC_PREP
        if (rc) return rc;
        ctx->sbound_seeded = true;
        if (record_events) QCK(cudaEventRecord(ctx->ev[2], ctx->stream));
        return QADC_OK;
    }
    // 2+3. residual, rotation, float tables
    {
        dim3 tgrid((ma + 7) / 8, nq);
        auto launch_tables = [&](auto kernel) {
            // 8 warps x (residual + rotated copy) x dim floats; above 48 KB (dim > 768, e.g. GIST-960) the kernel opts in
            if (static_cast<size_t>(dim) * 64 > 48 * 1024)
                cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dim * 64);
            kernel<<<tgrid, 256, static_cast<size_t>(dim) * 8 * 8, ctx->stream>>>(
                d_queries, dim, M, ctx->d_codebooks, ctx->d_rotation, flat ? nullptr : ctx->d_centroids, d_assign, ma,
                ctx->b_tables.as<float>(), ctx->b_tmin.as<float>(), 4, nullptr);
        };
        switch (dim / M) {   // same arithmetic in every instantiation; the common sub-vector sizes unroll
            case 2: launch_tables(tables_kernel<2>); break;
            case 3: launch_tables(tables_kernel<3>); break;
            case 4: launch_tables(tables_kernel<4>); break;
            case 6: launch_tables(tables_kernel<6>); break;
            case 8: launch_tables(tables_kernel<8>); break;
            case 12: launch_tables(tables_kernel<12>); break;
            case 16: launch_tables(tables_kernel<16>); break;
            default: launch_tables(tables_kernel<0>); break;
        }
        ctx->launches++;
        QCK(cudaGetLastError());
    }
    if (record_events) QCK(cudaEventRecord(ctx->ev[2], ctx->stream));
    // 4'. flat database with a long prefix: int8 lower bounds exclude all but a few thousand prefix vectors per query
    //     from the float scan (qadc_flatprep.cuh); qmax and the int8 tables are bit-identical to the plain path below
    if (flat && ctx->d_starts_native && ctx->opt_flat_prep && r <= kPrefMaxR) {
        const uint32_t n_prefix = ctx->h_start_size[0];
        const size_t te = static_cast<size_t>(M) * 16;
        const int nsplit = static_cast<int>(std::min(n_prefix, kPrefSampleMax) / kPrefSplit);
        ENSURE(ctx->b_prov, static_cast<size_t>(nq) * te);
        ENSURE(ctx->b_ghist, static_cast<size_t>(nq) * 129 * 4);   // [nq][128] the scan's global histograms | [nq] candidate counts
        // candidate slots per query: 16 384, fewer for large batches (256 MB in all, never below the 4096 the one-buffer
        // path of the final kernel takes; a query with more candidates than slots evaluates its whole prefix)
        const uint32_t cand_cap = static_cast<uint32_t>(std::max<size_t>(kPrefFastCap, std::min<size_t>(kPrefCandCap, (size_t(64) << 20) / nq)));
        ENSURE(ctx->b_cand, static_cast<size_t>(nq) * cand_cap * 4);
        ENSURE(ctx->b_plists, static_cast<size_t>(nq) * nsplit * r * 4);
        ENSURE(ctx->b_sbound, static_cast<size_t>(nq) * 4);
        QCK(cudaMemsetAsync(ctx->b_ghist.p, 0, static_cast<size_t>(nq) * 129 * 4, ctx->stream));   // one memset for both
        ctx->ghist_zeroed = true;
        FlatPrepArgs fa;
        fa.starts = ctx->d_starts; fa.native = ctx->d_starts_native; fa.n_prefix = n_prefix;
        fa.tables = ctx->b_tables.as<float>(); fa.r = r; fa.nsplit = nsplit; fa.sample_lists = ctx->b_plists.as<uint32_t>();
        fa.prov_qt = ctx->b_prov.as<int8_t>(); fa.cand_count = ctx->b_ghist.as<unsigned int>() + static_cast<size_t>(nq) * 128;
        fa.cand = ctx->b_cand.as<uint32_t>(); fa.cand_cap = cand_cap; fa.seed_out = ctx->b_sbound.as<int>();
        const uint32_t n_sb = (n_prefix + kSbVec - 1) / kSbVec;
        dim3 sgrid(nsplit, nq), pgrid(flat_prep_splits(ctx, n_sb, nq), nq);
        const PipeK pk = make_pipek();
        float* tabs = ctx->b_tables.as<float>();
        if (M == 16) {
            flat_prefix_sample_kernel<16><<<sgrid, kSelThreads, 0, ctx->stream>>>(fa);
            flat_prefix_bound_kernel<16><<<nq, kSelThreads, 0, ctx->stream>>>(fa);
            flat_prefix_pass_kernel<16, 1><<<pgrid, 256, 0, ctx->stream>>>(fa.native, n_prefix, fa.prov_qt, nullptr, nullptr, nullptr, 127, r, 0, fa.cand_count, fa.cand, cand_cap, pk);
            flat_bounds_final_kernel<16><<<nq, kSelThreads, 0, ctx->stream>>>(fa, tabs, ctx->b_tmin.as<float>(), ctx->b_qtables.as<int8_t>(),
                                                                             ctx->b_qmin.as<float>(), ctx->b_qmax.as<float>(), ctx->d_err);
        } else {
            flat_prefix_sample_kernel<32><<<sgrid, kSelThreads, 0, ctx->stream>>>(fa);
            flat_prefix_bound_kernel<32><<<nq, kSelThreads, 0, ctx->stream>>>(fa);
            flat_prefix_pass_kernel<32, 1><<<pgrid, 256, 0, ctx->stream>>>(fa.native, n_prefix, fa.prov_qt, nullptr, nullptr, nullptr, 127, r, 0, fa.cand_count, fa.cand, cand_cap, pk);
            flat_bounds_final_kernel<32><<<nq, kSelThreads, 0, ctx->stream>>>(fa, tabs, ctx->b_tmin.as<float>(), ctx->b_qtables.as<int8_t>(),
                                                                             ctx->b_qmin.as<float>(), ctx->b_qmax.as<float>(), ctx->d_err);
        }
        ctx->launches += 4;
        QCK(cudaGetLastError());
        ctx->sbound_seeded = ctx->opt_flat_seed != 0;   // the final kernel left the scan's shared-bound seed of this batch
        return QADC_OK;
    }
    // 4. keep-prefix float scan -> qmax
    PrefixArgs pa;
    pa.starts = ctx->d_starts; pa.start_off = ctx->d_start_off; pa.start_size = ctx->d_start_size;
    pa.assign = d_assign; pa.tables = ctx->b_tables.as<float>(); pa.ma = ma; pa.r = r; pa.M = M;
    pa.qmax = ctx->b_qmax.as<float>();
    const uint32_t* sel_lists = nullptr;
    int sel_n = 0;
    if (!flat && ctx->max_start <= 128) {
        // inverted lists with short prefixes: one warp per probe
        pa.nsplit = 1; pa.lists = nullptr;
        if (M == 16) prefix_scan_probes_kernel<16><<<nq, kSelThreads, 0, ctx->stream>>>(pa);
        else prefix_scan_probes_kernel<32><<<nq, kSelThreads, 0, ctx->stream>>>(pa);
        ctx->launches++;
        QCK(cudaGetLastError());
    } else {
        int nsplit = 1;
        if (flat) nsplit = static_cast<int>(std::min<uint32_t>(64, std::max<uint32_t>(1, ctx->max_start / 8192)));
        ENSURE(ctx->b_plists, static_cast<size_t>(nq) * nsplit * r * 8);
        pa.nsplit = nsplit; pa.lists = ctx->b_plists.as<uint32_t>();
        dim3 pgrid(nsplit, nq);
        if (M == 16) prefix_scan_kernel<16><<<pgrid, kSelThreads, 0, ctx->stream>>>(pa);
        else prefix_scan_kernel<32><<<pgrid, kSelThreads, 0, ctx->stream>>>(pa);
        ctx->launches++;
        QCK(cudaGetLastError());
        if (nsplit > 1) { sel_lists = pa.lists; sel_n = nsplit * r; }   // selected inside quantize_kernel
    }
    // 5. bounds + int8 tables
    quantize_kernel<<<nq, 256, 0, ctx->stream>>>(
        ctx->b_tables.as<float>(), ctx->b_tmin.as<float>(), ctx->b_qmax.as<float>(), ma, M,
        ctx->b_qtables.as<int8_t>(), ctx->b_qmin.as<float>(), ctx->d_err, nullptr, nullptr, nullptr,
        sel_lists, sel_n, r, ctx->b_qmax.as<float>());
    ctx->launches++;
    QCK(cudaGetLastError());
    return QADC_OK;
}

template <typename K>
int launch_tables_kernel(qadc_ctx* ctx, K kernel, dim3 grid, const float* d_queries, const int32_t* d_assign, int ma,
                                const uint32_t* part_size) {
    const int dim = ctx->dim;
    if (static_cast<size_t>(dim) * 64 > 48 * 1024)
        QCK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dim * 64));
    kernel<<<grid, 256, static_cast<size_t>(dim) * 8 * 8, ctx->stream>>>(d_queries, dim, ctx->m, ctx->d_codebooks, ctx->d_rotation,
                                                                        ctx->d_centroids, d_assign, ma, ctx->b_tables.as<float>(),
                                                                        ctx->b_tmin.as<float>(), 4, part_size);
    ctx->launches++;
    QCK(cudaGetLastError());
    return QADC_OK;
}

int check_search_args(qadc_ctx* ctx, int nq, int ma, int r) {
    if (!ctx) return QADC_EINVAL;
    if (!ctx->finalized) return fail(ctx, QADC_ESTATE, "database not finalized");
    if (nq <= 0 || ma <= 0 || r <= 0) return fail(ctx, QADC_EINVAL, "nq, ma and r must be positive");
    if (r > kMergeCap / 2) return fail(ctx, QADC_EINVAL, "r > 1024 is not supported");
    if (ctx->K == 0 && ma != 1)
        return fail(ctx, QADC_EINVAL, "a flat database must be queried with ma = 1 (SURVEY App. B)");
    if (ctx->K > 0 && (ma > ctx->K || ma > kSelCap / 2))
        return fail(ctx, QADC_EINVAL, "ma exceeds the partition count (or 1024)");
    return QADC_OK;
}

}  // namespace

// ============================================================================================
extern "C" {

int qadc_abi_version(void) { return QADC_ABI_VERSION; }

int qadc_create(int device, void* stream, qadc_ctx** out) {
    qadc_ctx* ctx = nullptr;
    if (!out) return fail(nullptr, QADC_EINVAL, "out is null");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, QADC_ECUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(nullptr, QADC_EINVAL, "bad device ordinal");
    cudaDeviceProp prop;
    QCK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(nullptr, QADC_ECUDA, "device is not sm_100 (libqadc_b200 has no other code path)");
    QCK(cudaSetDevice(device));
    qadc_ctx* c = new qadc_ctx;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    auto init = [&]() -> int {
        qadc_ctx* ctx = c;   // errors are reported on the new context, copied out below
        if (stream) c->stream = static_cast<cudaStream_t>(stream);
        else { QCK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
        for (auto& ev : c->ev) QCK(cudaEventCreate(&ev));
        for (auto& ev : c->ev_scan0) QCK(cudaEventCreate(&ev));
        for (auto& ev : c->ev_scan1) QCK(cudaEventCreate(&ev));
        QCK(cudaMalloc(&c->d_err, sizeof(int)));
        QCK(cudaMemset(c->d_err, 0, sizeof(int)));
        QCK(cudaMallocHost(&c->h_err, sizeof(int)));
        *c->h_err = 0;
        return QADC_OK;
    };
    const int rc = init();
    if (rc != QADC_OK) {
        g_create_error = c->err;
        qadc_destroy(c);
        return rc;
    }
    *out = c;
    return QADC_OK;
}

void qadc_destroy(qadc_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    free_db(c);
    cudaFree(c->d_codebooks); cudaFree(c->d_rotation); cudaFree(c->d_centroids);
    for (DevBuf* b : {&c->staging, &c->b_queries, &c->b_assign, &c->b_tables, &c->b_tmin, &c->b_qmax, &c->b_qmin,
                      &c->b_qtables, &c->b_lists, &c->b_plists, &c->b_ids, &c->b_dists, &c->b_counts, &c->b_keys,
                      &c->b_dump, &c->b_hist, &c->b_sbound, &c->b_cdist, &c->b_adc_dists, &c->b_qmin_raw, &c->b_prov,
                      &c->b_cand, &c->b_ghist})
        cudaFree(b->p);
    cudaFree(c->d_rows); cudaFree(c->d_row_off); cudaFree(c->d_adc_labels);
    cudaFree(c->d_err);
    if (c->h_err) cudaFreeHost(c->h_err);
    for (auto& ev : c->ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : c->ev_scan0) if (ev) cudaEventDestroy(ev);
    for (auto& ev : c->ev_scan1) if (ev) cudaEventDestroy(ev);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char* qadc_last_error(const qadc_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int qadc_set_pq(qadc_ctx* ctx, int dim, int m, int bits, const float* codebooks, const float* rotation) {
    if (!ctx || !codebooks) return fail(ctx, QADC_EINVAL, "null argument");
    // get_simd_scan_func_epi8 (db_query_4.cpp:23-35), load_database_check (:393-402)
    // plus the 8- and 16-bit configurations of the plain ADC tool (get_scan_func, query_common.hpp:122-147); those
    // contexts only serve qadc_adc_load / qadc_adc_search (and qadc_encode)
    const bool quick = bits == 4 && (m == 16 || m == 32);
    const bool adc8 = bits == 8 && (m == 4 || m == 8 || m == 16);
    const bool adc16 = bits == 16 && (m == 2 || m == 4 || m == 8);
    if (!quick && !adc8 && !adc16)
        return fail(ctx, QADC_EINVAL, "Unsupported (nsq,nsq_bits) configuration. Supported: (16,4) (32,4); plain ADC also (4,8) (8,8) (16,8) (2,16) (4,16) (8,16).");
    if (dim <= 0 || dim % m != 0) return fail(ctx, QADC_EINVAL, "dim must be a positive multiple of m");
    if (static_cast<size_t>(dim) * 64 > static_cast<size_t>(kMaxSmem))
        return fail(ctx, QADC_EINVAL, "dim > 3632 is not supported (the table kernel keeps 8 x 2 x dim floats in shared memory)");
    QCK(cudaSetDevice(ctx->device));
    cudaFree(ctx->d_codebooks); cudaFree(ctx->d_rotation);
    ctx->d_codebooks = nullptr; ctx->d_rotation = nullptr;
    const size_t cb = (static_cast<size_t>(dim) << bits) * sizeof(float);
    QCK(cudaMalloc(&ctx->d_codebooks, cb));
    QCK(cudaMemcpy(ctx->d_codebooks, codebooks, cb, cudaMemcpyHostToDevice));
    if (rotation) {
        const size_t rb = static_cast<size_t>(dim) * dim * sizeof(float);
        QCK(cudaMalloc(&ctx->d_rotation, rb));
        QCK(cudaMemcpy(ctx->d_rotation, rotation, rb, cudaMemcpyHostToDevice));
    }
    ctx->dim = dim; ctx->m = m; ctx->bits = bits;
    return QADC_OK;
}

int qadc_set_coarse(qadc_ctx* ctx, int K, const float* centroids) {
    if (!ctx) return QADC_EINVAL;
    if (ctx->dim == 0) return fail(ctx, QADC_ESTATE, "qadc_set_pq must be called first");
    QCK(cudaSetDevice(ctx->device));
    cudaFree(ctx->d_centroids); ctx->d_centroids = nullptr; ctx->K = 0;
    if (K <= 0) return QADC_OK;
    if (!centroids) return fail(ctx, QADC_EINVAL, "centroids is null");
    const size_t bytes = static_cast<size_t>(K) * ctx->dim * sizeof(float);
    QCK(cudaMalloc(&ctx->d_centroids, bytes));
    QCK(cudaMemcpy(ctx->d_centroids, centroids, bytes, cudaMemcpyHostToDevice));
    ctx->K = K;
    return QADC_OK;
}

int qadc_begin_database(qadc_ctx* ctx, int partition_count, const uint32_t* sizes, int has_labels) {
    if (!ctx || !sizes || partition_count <= 0) return fail(ctx, QADC_EINVAL, "bad database description");
    if (ctx->m == 0) return fail(ctx, QADC_ESTATE, "qadc_set_pq must be called first");
    if (ctx->bits != 4) return fail(ctx, QADC_EINVAL, "Quantizer must have sq_bits=4");   // load_database_check, db_query_4.cpp:393-402
    if (ctx->K == 0 && partition_count != 1) return fail(ctx, QADC_EINVAL, "a flat database has exactly one partition");
    if (ctx->K > 0 && partition_count != ctx->K) return fail(ctx, QADC_EINVAL, "partition_count != coarse centroid count");
    QCK(cudaSetDevice(ctx->device));
    free_db(ctx);
    const int P = partition_count;
    ctx->parts = P; ctx->has_labels = has_labels != 0;
    ctx->h_size.assign(sizes, sizes + P);
    ctx->h_pos_base.assign(P, 0);
    ctx->h_sb_off.assign(P, 0); ctx->h_label_off.assign(P, 0);
    ctx->h_explicit_prefix.assign(P, nullptr); ctx->h_explicit_count.assign(P, 0);
    uint64_t sb = 0, vec = 0;
    for (int p = 0; p < P; ++p) {
        ctx->h_sb_off[p] = sb; ctx->h_label_off[p] = vec;
        sb += (static_cast<uint64_t>(sizes[p]) + kSbVec - 1) / kSbVec;
        vec += sizes[p];
    }
    ctx->total_sb = sb; ctx->total_vec = vec;
    QCK(cudaMalloc(&ctx->d_codes, std::max<size_t>(16, sb * sb_bytes(ctx->m))));
    if (ctx->has_labels) QCK(cudaMalloc(&ctx->d_labels, std::max<size_t>(4, vec * 4)));
    QCK(cudaMalloc(&ctx->d_sb_off, P * 8)); QCK(cudaMalloc(&ctx->d_label_off, P * 8));
    QCK(cudaMalloc(&ctx->d_size, P * 4));
    QCK(cudaMemcpy(ctx->d_sb_off, ctx->h_sb_off.data(), P * 8, cudaMemcpyHostToDevice));
    QCK(cudaMemcpy(ctx->d_label_off, ctx->h_label_off.data(), P * 8, cudaMemcpyHostToDevice));
    QCK(cudaMemcpy(ctx->d_size, ctx->h_size.data(), P * 4, cudaMemcpyHostToDevice));
    {   // default ownership: the partitions that hold vectors here
        std::vector<uint32_t> owned(P);
        for (int p = 0; p < P; ++p) owned[p] = sizes[p] ? 1u : 0u;
        QCK(cudaMalloc(&ctx->d_owned, P * 4));
        QCK(cudaMemcpy(ctx->d_owned, owned.data(), P * 4, cudaMemcpyHostToDevice));
    }
    ctx->begun = true;
    return QADC_OK;
}

int qadc_set_owned_partitions(qadc_ctx* ctx, const uint8_t* owned) {
    if (!ctx || !ctx->begun) return fail(ctx, QADC_ESTATE, "qadc_begin_database must be called first");
    if (!owned) return fail(ctx, QADC_EINVAL, "null mask");
    QCK(cudaSetDevice(ctx->device));
    std::vector<uint32_t> o(ctx->parts);
    for (int p = 0; p < ctx->parts; ++p) {
        if (!owned[p] && ctx->h_size[p]) return fail(ctx, QADC_EINVAL, "a partition that holds vectors on this shard must be owned by it");
        o[p] = owned[p] ? 1u : 0u;
    }
    QCK(cudaStreamSynchronize(ctx->stream));
    QCK(cudaMemcpy(ctx->d_owned, o.data(), o.size() * 4, cudaMemcpyHostToDevice));
    return QADC_OK;
}

int qadc_upload_codes(qadc_ctx* ctx, int part_i, uint32_t first, uint32_t count, const uint8_t* codes,
                      const uint32_t* labels, int on_device) {
    if (!ctx || !ctx->begun) return fail(ctx, QADC_ESTATE, "qadc_begin_database must be called first");
    if (part_i < 0 || part_i >= ctx->parts) return fail(ctx, QADC_EINVAL, "bad partition index");
    if (count == 0) return QADC_OK;
    if (!codes) return fail(ctx, QADC_EINVAL, "codes is null");
    if (first % kSbVec != 0) return fail(ctx, QADC_EINVAL, "first must be a multiple of 256");
    if (static_cast<uint64_t>(first) + count > ctx->h_size[part_i]) return fail(ctx, QADC_EINVAL, "range exceeds partition size");
    if (first + count != ctx->h_size[part_i] && count % kSbVec != 0)
        return fail(ctx, QADC_EINVAL, "only the last chunk of a partition may be ragged");
    if (ctx->has_labels && !labels) return fail(ctx, QADC_EINVAL, "labels is null but the database has labels");
    QCK(cudaSetDevice(ctx->device));
    const int M = ctx->m, CS = M / 2;
    uint8_t* dst = ctx->d_codes + ctx->h_sb_off[part_i] * sb_bytes(M);
    const uint32_t kChunk = 1u << 23;   // vectors per staging chunk (64/128 MiB)
    for (uint32_t off = 0; off < count; off += kChunk) {
        const uint32_t n = std::min(kChunk, count - off);
        const uint8_t* src = codes + static_cast<size_t>(off) * CS;
        if (!on_device) {
            ENSURE(ctx->staging, static_cast<size_t>(n) * CS);
            QCK(cudaMemcpyAsync(ctx->staging.p, src, static_cast<size_t>(n) * CS, cudaMemcpyHostToDevice, ctx->stream));
            src = ctx->staging.as<uint8_t>();
        }
        const uint32_t groups = ((n + kSbVec - 1) / kSbVec) * 32;
        const unsigned blocks = (groups + 255) / 256;
        if (M == 16) transpose_codes_kernel<16><<<blocks, 256, 0, ctx->stream>>>(src, n, dst, first + off);
        else transpose_codes_kernel<32><<<blocks, 256, 0, ctx->stream>>>(src, n, dst, first + off);
        QCK(cudaGetLastError());
        if (!on_device) QCK(cudaStreamSynchronize(ctx->stream));   // staging is reused
    }
    if (ctx->has_labels)
        QCK(cudaMemcpyAsync(ctx->d_labels + ctx->h_label_off[part_i] + first, labels, static_cast<size_t>(count) * 4,
                            on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
    QCK(cudaStreamSynchronize(ctx->stream));
    ctx->finalized = false;
    return QADC_OK;
}

// Uploads the partitions [p0, p1) from one contiguous source (device pointer, or host through the staging buffer).
static int upload_run(qadc_ctx* ctx, int p0, int p1, const uint8_t* codes, const uint32_t* labels, bool src_on_device) {
    const int M = ctx->m, CS = M / 2;
    const uint64_t v0 = ctx->h_label_off[p0];
    const uint64_t v1 = (p1 < ctx->parts) ? ctx->h_label_off[p1] : ctx->total_vec;
    const uint64_t sb0 = ctx->h_sb_off[p0], sb1 = (p1 < ctx->parts) ? ctx->h_sb_off[p1] : ctx->total_sb;
    if (v1 == v0) return QADC_OK;
    const uint8_t* d_src = codes;
    if (!src_on_device) {
        ENSURE(ctx->staging, (v1 - v0) * CS);
        QCK(cudaMemcpyAsync(ctx->staging.p, codes, (v1 - v0) * CS, cudaMemcpyHostToDevice, ctx->stream));
        d_src = ctx->staging.as<uint8_t>();
    }
    const uint64_t threads = (sb1 - sb0) * 32;
    const unsigned blocks = static_cast<unsigned>((threads + 255) / 256);
    if (M == 16)
        transpose_partitions_kernel<16><<<blocks, 256, 0, ctx->stream>>>(d_src, p0, p1, ctx->d_sb_off, ctx->d_label_off, ctx->d_size, sb1, ctx->d_codes);
    else
        transpose_partitions_kernel<32><<<blocks, 256, 0, ctx->stream>>>(d_src, p0, p1, ctx->d_sb_off, ctx->d_label_off, ctx->d_size, sb1, ctx->d_codes);
    QCK(cudaGetLastError());
    if (ctx->has_labels)
        QCK(cudaMemcpyAsync(ctx->d_labels + v0, labels, (v1 - v0) * 4, src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                            ctx->stream));
    QCK(cudaStreamSynchronize(ctx->stream));   // the staging buffer / the caller's buffers are free again
    return QADC_OK;
}

int qadc_upload_database(qadc_ctx* ctx, const uint8_t* codes, const uint32_t* labels, int on_device) {
    if (!ctx || !ctx->begun) return fail(ctx, QADC_ESTATE, "qadc_begin_database must be called first");
    if (ctx->total_vec == 0) return QADC_OK;
    if (!codes) return fail(ctx, QADC_EINVAL, "codes is null");
    if (ctx->has_labels && !labels) return fail(ctx, QADC_EINVAL, "labels is null but the database has labels");
    QCK(cudaSetDevice(ctx->device));
    const int CS = ctx->m / 2;
    const uint64_t kChunk = 1u << 23;   // vectors per run (64 / 128 MiB of staging)
    int p0 = 0;
    while (p0 < ctx->parts) {
        const uint64_t v0 = ctx->h_label_off[p0];
        if (ctx->h_size[p0] > kChunk) {   // a partition larger than a run goes through the chunked single-partition path
            int rc = qadc_upload_codes(ctx, p0, 0, ctx->h_size[p0], codes + v0 * CS, labels ? labels + v0 : nullptr, on_device);
            if (rc) return rc;
            ++p0;
            continue;
        }
        int p1 = p0 + 1;
        while (p1 < ctx->parts && ctx->h_size[p1] <= kChunk && ctx->h_label_off[p1] + ctx->h_size[p1] - v0 <= kChunk) ++p1;
        int rc = upload_run(ctx, p0, p1, codes + v0 * CS, labels ? labels + v0 : nullptr, on_device != 0);
        if (rc) return rc;
        p0 = p1;
    }
    ctx->finalized = false;
    return QADC_OK;
}

int qadc_upload_partitions(qadc_ctx* ctx, const uint8_t* const* part_codes, const uint32_t* const* part_labels) {
    if (!ctx || !ctx->begun) return fail(ctx, QADC_ESTATE, "qadc_begin_database must be called first");
    if (!part_codes) return fail(ctx, QADC_EINVAL, "part_codes is null");
    if (ctx->has_labels && !part_labels) return fail(ctx, QADC_EINVAL, "part_labels is null but the database has labels");
    QCK(cudaSetDevice(ctx->device));
    const size_t CS = ctx->m / 2;
    const uint64_t kChunk = 1u << 22;
    std::vector<uint8_t> hc;
    std::vector<uint32_t> hl;
    int p0 = 0;
    while (p0 < ctx->parts) {
        if (ctx->h_size[p0] > kChunk) {
            int rc = qadc_upload_codes(ctx, p0, 0, ctx->h_size[p0], part_codes[p0], ctx->has_labels ? part_labels[p0] : nullptr, 0);
            if (rc) return rc;
            ++p0;
            continue;
        }
        // gather a run of short partitions into one host buffer, then one copy + one re-layout launch
        int p1 = p0;
        uint64_t n = 0;
        while (p1 < ctx->parts && ctx->h_size[p1] <= kChunk && n + ctx->h_size[p1] <= kChunk) { n += ctx->h_size[p1]; ++p1; }
        hc.resize(n * CS);
        if (ctx->has_labels) hl.resize(n);
        uint64_t at = 0;
        for (int p = p0; p < p1; ++p) {
            const uint32_t sz = ctx->h_size[p];
            if (!sz) continue;
            if (!part_codes[p] || (ctx->has_labels && !part_labels[p])) return fail(ctx, QADC_EINVAL, "null partition pointer");
            memcpy(hc.data() + at * CS, part_codes[p], sz * CS);
            if (ctx->has_labels) memcpy(hl.data() + at, part_labels[p], static_cast<size_t>(sz) * 4);
            at += sz;
        }
        int rc = upload_run(ctx, p0, p1, hc.data(), ctx->has_labels ? hl.data() : nullptr, false);
        if (rc) return rc;
        p0 = p1;
    }
    ctx->finalized = false;
    return QADC_OK;
}

int qadc_set_position_base(qadc_ctx* ctx, int part_i, uint32_t pos_base) {
    if (!ctx || !ctx->begun) return fail(ctx, QADC_ESTATE, "qadc_begin_database must be called first");
    if (part_i < 0 || part_i >= ctx->parts) return fail(ctx, QADC_EINVAL, "bad partition index");
    ctx->h_pos_base[part_i] = pos_base;
    ctx->finalized = false;
    return QADC_OK;
}

int qadc_set_prefix(qadc_ctx* ctx, int part_i, const uint8_t* codes, uint32_t count, int on_device) {
    if (!ctx || !ctx->begun) return fail(ctx, QADC_ESTATE, "qadc_begin_database must be called first");
    if (part_i < 0 || part_i >= ctx->parts) return fail(ctx, QADC_EINVAL, "bad partition index");
    if (!codes || count == 0) return fail(ctx, QADC_EINVAL, "empty prefix");
    QCK(cudaSetDevice(ctx->device));
    if (ctx->d_prefix_block) return fail(ctx, QADC_ESTATE, "prefixes were set in bulk (qadc_set_prefixes)");
    cudaFree(ctx->h_explicit_prefix[part_i]);
    ctx->h_explicit_prefix[part_i] = nullptr;
    const size_t bytes = static_cast<size_t>(count) * (ctx->m / 2);
    QCK(cudaMalloc(&ctx->h_explicit_prefix[part_i], bytes));
    QCK(cudaMemcpy(ctx->h_explicit_prefix[part_i], codes, bytes, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
    ctx->h_explicit_count[part_i] = count;
    ctx->finalized = false;
    return QADC_OK;
}

int qadc_set_prefixes(qadc_ctx* ctx, const uint8_t* codes, const uint32_t* counts, int on_device) {
    if (!ctx || !ctx->begun) return fail(ctx, QADC_ESTATE, "qadc_begin_database must be called first");
    if (!codes || !counts) return fail(ctx, QADC_EINVAL, "null argument");
    QCK(cudaSetDevice(ctx->device));
    for (auto& p : ctx->h_explicit_prefix) {
        if (p && !ctx->d_prefix_block) cudaFree(p);
        p = nullptr;
    }
    cudaFree(ctx->d_prefix_block); ctx->d_prefix_block = nullptr;
    const size_t CS = ctx->m / 2;
    uint64_t total = 0;
    for (int p = 0; p < ctx->parts; ++p) total += counts[p];
    if (total == 0) return fail(ctx, QADC_EINVAL, "empty prefixes");
    QCK(cudaMalloc(&ctx->d_prefix_block, total * CS));
    QCK(cudaMemcpy(ctx->d_prefix_block, codes, total * CS, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
    uint64_t at = 0;
    for (int p = 0; p < ctx->parts; ++p) {
        ctx->h_explicit_prefix[p] = counts[p] ? ctx->d_prefix_block + at * CS : nullptr;
        ctx->h_explicit_count[p] = counts[p];
        at += counts[p];
    }
    ctx->finalized = false;
    return QADC_OK;
}

int qadc_finalize(qadc_ctx* ctx, float keep) {
    if (!ctx || !ctx->begun) return fail(ctx, QADC_ESTATE, "qadc_begin_database must be called first");
    QCK(cudaSetDevice(ctx->device));
    const int P = ctx->parts, M = ctx->m, CS = M / 2;
    ctx->h_start_size.assign(P, 0); ctx->h_start_off.assign(P, 0);
    std::vector<uint8_t> has_explicit(P, 0);
    uint64_t off = 0;
    ctx->max_start = 0;
    for (int p = 0; p < P; ++p) {
        uint32_t s;
        if (ctx->h_explicit_prefix[p]) { s = ctx->h_explicit_count[p]; has_explicit[p] = 1; }
        else {
            if (ctx->h_pos_base[p] != 0)
                return fail(ctx, QADC_EINVAL, "a shard with a position base needs an explicit prefix (qadc_set_prefix)");
            const uint32_t size = ctx->h_size[p];
            // starts_sizes[p] = max(1u, (unsigned)(size * keep)), float32 (db_query_4.cpp:125-126)
            s = size == 0 ? 0u : std::max(1u, static_cast<unsigned>(static_cast<float>(size) * keep));
            s = std::min(s, size);
        }
        ctx->h_start_size[p] = s; ctx->h_start_off[p] = off; off += s;
        ctx->max_start = std::max(ctx->max_start, s);
    }
    cudaFree(ctx->d_starts); cudaFree(ctx->d_start_off); cudaFree(ctx->d_start_size); cudaFree(ctx->d_pos_base);
    ctx->d_starts = nullptr; ctx->d_start_off = nullptr; ctx->d_start_size = nullptr; ctx->d_pos_base = nullptr;
    QCK(cudaMalloc(&ctx->d_starts, std::max<size_t>(16, off * CS)));
    QCK(cudaMalloc(&ctx->d_start_off, P * 8)); QCK(cudaMalloc(&ctx->d_start_size, P * 4));
    QCK(cudaMalloc(&ctx->d_pos_base, P * 4));
    QCK(cudaMemcpy(ctx->d_start_off, ctx->h_start_off.data(), P * 8, cudaMemcpyHostToDevice));
    QCK(cudaMemcpy(ctx->d_start_size, ctx->h_start_size.data(), P * 4, cudaMemcpyHostToDevice));
    QCK(cudaMemcpy(ctx->d_pos_base, ctx->h_pos_base.data(), P * 4, cudaMemcpyHostToDevice));
    ENSURE(ctx->b_dump, static_cast<size_t>(P));
    uint8_t* d_flags = ctx->b_dump.as<uint8_t>();
    QCK(cudaMemcpy(d_flags, has_explicit.data(), P, cudaMemcpyHostToDevice));
    dim3 grid(P, std::max(1u, std::min(64u, (ctx->max_start + 255) / 256)));
    if (M == 16)
        extract_prefix_kernel<16><<<grid, 256, 0, ctx->stream>>>(ctx->d_codes, ctx->d_sb_off, ctx->d_start_size,
                                                                 ctx->d_start_off, d_flags, ctx->d_starts);
    else
        extract_prefix_kernel<32><<<grid, 256, 0, ctx->stream>>>(ctx->d_codes, ctx->d_sb_off, ctx->d_start_size,
                                                                 ctx->d_start_off, d_flags, ctx->d_starts);
    QCK(cudaGetLastError());
    bool all_explicit = ctx->d_prefix_block != nullptr;
    for (int p = 0; p < P && all_explicit; ++p) all_explicit = ctx->h_explicit_prefix[p] || ctx->h_start_size[p] == 0;
    if (all_explicit) {
        // set in bulk for every partition: the block has the layout of d_starts, one copy
        QCK(cudaMemcpyAsync(ctx->d_starts, ctx->d_prefix_block, off * CS, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
        for (int p = 0; p < P; ++p)
            if (ctx->h_explicit_prefix[p])
                QCK(cudaMemcpyAsync(ctx->d_starts + ctx->h_start_off[p] * CS, ctx->h_explicit_prefix[p],
                                    static_cast<size_t>(ctx->h_explicit_count[p]) * CS, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    cudaFree(ctx->d_starts_native); ctx->d_starts_native = nullptr;
    if (ctx->K == 0 && ctx->h_start_size[0] >= kPrefMinNative) {
        // flat database with a long keep-prefix: a nibble-plane copy of the prefix for the int8 passes of qadc_flatprep.cuh
        const uint32_t n = ctx->h_start_size[0];
        const size_t n_sb = (static_cast<size_t>(n) + kSbVec - 1) / kSbVec;
        QCK(cudaMalloc(&ctx->d_starts_native, n_sb * sb_bytes(M)));
        const unsigned blocks = static_cast<unsigned>((n_sb * 32 + 255) / 256);
        if (M == 16) transpose_codes_kernel<16><<<blocks, 256, 0, ctx->stream>>>(ctx->d_starts, n, ctx->d_starts_native, 0);
        else transpose_codes_kernel<32><<<blocks, 256, 0, ctx->stream>>>(ctx->d_starts, n, ctx->d_starts_native, 0);
        QCK(cudaGetLastError());
    }
    QCK(cudaStreamSynchronize(ctx->stream));
    ctx->finalized = true;
    return QADC_OK;
}

static int search_device_impl(qadc_ctx* ctx, const float* d_queries, const int32_t* d_assign_in, int nq, int ma, int r,
                              uint32_t* d_ids, int8_t* d_dists, int32_t* d_counts, uint64_t* d_keys) {
    int rc = check_search_args(ctx, nq, ma, r);
    if (rc) return rc;
    QCK(cudaSetDevice(ctx->device));
    ctx->launches = 0;
    // sub-batches keep every grid dimension below 65 536 and the per-batch scratch bounded
    for (int q0 = 0; q0 < nq; q0 += kMaxBatch) {
        const int n = std::min(kMaxBatch, nq - q0);
        rc = tables_device(ctx, d_queries + static_cast<size_t>(q0) * ctx->dim, n, ma, r,
                           d_assign_in ? d_assign_in + static_cast<size_t>(q0) * ma : nullptr, q0 == 0);
        if (rc) return rc;
        if (q0 == 0) QCK(cudaEventRecord(ctx->ev[3], ctx->stream));
        rc = scan_device(ctx, ctx->b_assign.as<int32_t>(), ctx->b_qtables.as<int8_t>(), n, ma, r,
                         d_ids + static_cast<size_t>(q0) * r, d_dists + static_cast<size_t>(q0) * r, d_counts + q0,
                         d_keys ? d_keys + static_cast<size_t>(q0) * r : nullptr);
        if (rc) return rc;
    }
    QCK(cudaEventRecord(ctx->ev[4], ctx->stream));
    QCK(cudaMemcpyAsync(ctx->h_err, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    return QADC_OK;
}

int qadc_search_device(qadc_ctx* ctx, const float* d_queries, int nq, int ma, int r, uint32_t* d_ids,
                       int8_t* d_dists, int32_t* d_counts, uint64_t* d_keys) {
    return search_device_impl(ctx, d_queries, nullptr, nq, ma, r, d_ids, d_dists, d_counts, d_keys);
}

int qadc_search_assigned_device(qadc_ctx* ctx, const float* d_queries, const int32_t* d_assign, int nq, int ma, int r,
                                uint32_t* d_ids, int8_t* d_dists, int32_t* d_counts, uint64_t* d_keys) {
    if (!ctx) return QADC_EINVAL;
    if (!d_assign) return fail(ctx, QADC_EINVAL, "null assignment");
    return search_device_impl(ctx, d_queries, d_assign, nq, ma, r, d_ids, d_dists, d_counts, d_keys);
}

// ---- "owner computes": the table pipeline of sharded inverted lists, split around one exchange -------------------
int qadc_tables_local_device(qadc_ctx* ctx, const float* d_queries, const int32_t* d_assign, int nq, int ma, int r,
                             float* d_local) {
    int rc = check_search_args(ctx, nq, ma, r);
    if (rc) return rc;
    if (ctx->K == 0) return fail(ctx, QADC_ESTATE, "qadc_tables_local_device needs inverted lists");
    if (!d_queries || !d_assign || !d_local) return fail(ctx, QADC_EINVAL, "null buffer");
    if (nq > kMaxBatch) return fail(ctx, QADC_EINVAL, "more than 32768 queries per call: split the batch");
    QCK(cudaSetDevice(ctx->device));
    ctx->launches = 0;
    ctx->local_nq = 0;
    const int M = ctx->m;
    const size_t nqa = static_cast<size_t>(nq) * ma;
    ENSURE(ctx->b_assign, nqa * 4);
    ENSURE(ctx->b_tables, nqa * M * 16 * 4);
    ENSURE(ctx->b_tmin, nqa * 4);
    ENSURE(ctx->b_qmax, static_cast<size_t>(nq) * 4);
    ENSURE(ctx->b_qmin, static_cast<size_t>(nq) * 4);
    ENSURE(ctx->b_qmin_raw, static_cast<size_t>(nq) * 4);
    ENSURE(ctx->b_qtables, nqa * M * 16);
    int32_t* assign = ctx->b_assign.as<int32_t>();
    QCK(cudaEventRecord(ctx->ev[0], ctx->stream));
    QCK(cudaMemcpyAsync(assign, d_assign, nqa * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    QCK(cudaEventRecord(ctx->ev[1], ctx->stream));
    const dim3 tgrid((ma + 7) / 8, nq);
    switch (ctx->dim / M) {
        case 2: rc = launch_tables_kernel(ctx, tables_kernel<2>, tgrid, d_queries, assign, ma, ctx->d_owned); break;
        case 3: rc = launch_tables_kernel(ctx, tables_kernel<3>, tgrid, d_queries, assign, ma, ctx->d_owned); break;
        case 4: rc = launch_tables_kernel(ctx, tables_kernel<4>, tgrid, d_queries, assign, ma, ctx->d_owned); break;
        case 6: rc = launch_tables_kernel(ctx, tables_kernel<6>, tgrid, d_queries, assign, ma, ctx->d_owned); break;
        case 8: rc = launch_tables_kernel(ctx, tables_kernel<8>, tgrid, d_queries, assign, ma, ctx->d_owned); break;
        case 12: rc = launch_tables_kernel(ctx, tables_kernel<12>, tgrid, d_queries, assign, ma, ctx->d_owned); break;
        case 16: rc = launch_tables_kernel(ctx, tables_kernel<16>, tgrid, d_queries, assign, ma, ctx->d_owned); break;
        default: rc = launch_tables_kernel(ctx, tables_kernel<0>, tgrid, d_queries, assign, ma, ctx->d_owned); break;
    }
    if (rc) return rc;
    PrefixArgs pa;
    pa.starts = ctx->d_starts; pa.start_off = ctx->d_start_off; pa.start_size = ctx->d_start_size;
    pa.assign = assign; pa.tables = ctx->b_tables.as<float>(); pa.ma = ma; pa.r = r; pa.M = M;
    pa.qmax = nullptr; pa.nsplit = 1; pa.lists = nullptr;
    pa.local_out = d_local; pa.tmin = ctx->b_tmin.as<float>(); pa.owned_size = ctx->d_owned;
    if (ctx->max_start <= 128) {
        if (M == 16) prefix_scan_probes_kernel<16><<<nq, kSelThreads, 0, ctx->stream>>>(pa);
        else prefix_scan_probes_kernel<32><<<nq, kSelThreads, 0, ctx->stream>>>(pa);
    } else {
        if (M == 16) prefix_scan_kernel<16><<<dim3(1, nq), kSelThreads, 0, ctx->stream>>>(pa);
        else prefix_scan_kernel<32><<<dim3(1, nq), kSelThreads, 0, ctx->stream>>>(pa);
    }
    ctx->launches++;
    QCK(cudaGetLastError());
    QCK(cudaEventRecord(ctx->ev[2], ctx->stream));
    ctx->local_nq = nq; ctx->local_ma = ma; ctx->local_r = r;
    return QADC_OK;
}

int qadc_search_bounded_device(qadc_ctx* ctx, const float* d_gathered, int G, int nq, int ma, int r, uint32_t* d_ids,
                               int8_t* d_dists, int32_t* d_counts, uint64_t* d_keys) {
    if (!ctx) return QADC_EINVAL;
    if (ctx->local_nq != nq || ctx->local_ma != ma || ctx->local_r != r || nq <= 0)
        return fail(ctx, QADC_ESTATE, "qadc_search_bounded_device must follow qadc_tables_local_device of the same batch");
    if (!d_gathered || !d_ids || !d_dists || !d_counts || G <= 0) return fail(ctx, QADC_EINVAL, "bad arguments");
    const size_t csmem = static_cast<size_t>(G) * r * 4;
    if (csmem > static_cast<size_t>(kMaxSmem) - 4096) return fail(ctx, QADC_EINVAL, "G * r too large");
    QCK(cudaSetDevice(ctx->device));
    ctx->local_nq = 0;
    const int M = ctx->m;
    if (csmem > 40 * 1024)
        QCK(cudaFuncSetAttribute(bounds_combine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(csmem)));
    bounds_combine_kernel<<<nq, kSelThreads, csmem, ctx->stream>>>(d_gathered, G, nq, r, ctx->b_qmin_raw.as<float>(),
                                                                  ctx->b_qmax.as<float>());
    QCK(cudaGetLastError());
    quantize_kernel<<<nq, 256, 0, ctx->stream>>>(ctx->b_tables.as<float>(), ctx->b_tmin.as<float>(), ctx->b_qmax.as<float>(), ma, M,
                                                 ctx->b_qtables.as<int8_t>(), ctx->b_qmin.as<float>(), ctx->d_err,
                                                 ctx->b_qmin_raw.as<float>(), ctx->b_assign.as<int32_t>(), ctx->d_owned);
    QCK(cudaGetLastError());
    ctx->launches += 2;
    QCK(cudaEventRecord(ctx->ev[3], ctx->stream));
    ctx->local_mode = true;
    const int rc = scan_device(ctx, ctx->b_assign.as<int32_t>(), ctx->b_qtables.as<int8_t>(), nq, ma, r, d_ids, d_dists, d_counts,
                               d_keys);
    ctx->local_mode = false;
    if (rc) return rc;
    QCK(cudaEventRecord(ctx->ev[4], ctx->stream));
    QCK(cudaMemcpyAsync(ctx->h_err, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    return QADC_OK;
}

int qadc_coarse_partial_device(qadc_ctx* ctx, const float* d_queries, int nq, int ma, int c_first, int c_count,
                               uint64_t* d_out_keys) {
    if (!ctx) return QADC_EINVAL;
    if (ctx->K == 0) return fail(ctx, QADC_ESTATE, "no coarse quantizer (flat database)");
    if (!d_queries || !d_out_keys) return fail(ctx, QADC_EINVAL, "null buffer");
    if (nq <= 0 || ma <= 0 || ma > kSelCap / 2 || c_first < 0 || c_count < 0 || c_first + c_count > ctx->K)
        return fail(ctx, QADC_EINVAL, "bad coarse range or ma");
    QCK(cudaSetDevice(ctx->device));
    const int dim = ctx->dim;
    if (c_count == 0) {
        QCK(cudaMemsetAsync(d_out_keys, 0xff, static_cast<size_t>(nq) * ma * 8, ctx->stream));
        return QADC_OK;
    }
    const int chunk_q = std::max(1, static_cast<int>(std::min<size_t>(nq, (size_t(1) << 28) / c_count)));
    ENSURE(ctx->b_cdist, static_cast<size_t>(chunk_q) * c_count * 4);
    for (int q0 = 0; q0 < nq; q0 += chunk_q) {
        const int n = std::min(chunk_q, nq - q0);
        dim3 grid((n + kCoarseTQ - 1) / kCoarseTQ, (c_count + kCoarseTC - 1) / kCoarseTC);
        coarse_dist_kernel<<<grid, 256, 0, ctx->stream>>>(d_queries + static_cast<size_t>(q0) * dim, n, dim,
                                                          ctx->d_centroids + static_cast<size_t>(c_first) * dim, c_count,
                                                          ctx->b_cdist.as<float>());
        QCK(cudaGetLastError());
        coarse_select_kernel<<<n, kSelThreads, 0, ctx->stream>>>(ctx->b_cdist.as<float>(), c_count, ma, nullptr,
                                                                d_out_keys + static_cast<size_t>(q0) * ma,
                                                                static_cast<uint32_t>(c_first));
        QCK(cudaGetLastError());
        ctx->launches += 2;
    }
    return QADC_OK;
}

int qadc_coarse_merge_device(qadc_ctx* ctx, const uint64_t* d_keys, int G, int nq, int ma, int32_t* d_assign) {
    if (!ctx) return QADC_EINVAL;
    if (!d_keys || !d_assign) return fail(ctx, QADC_EINVAL, "null buffer");
    if (G <= 0 || nq <= 0 || ma <= 0) return fail(ctx, QADC_EINVAL, "bad merge shape");
    const int n_sort = next_pow2(std::max(G * ma, 2));
    if (static_cast<size_t>(n_sort) * 8 > kMaxSmem) return fail(ctx, QADC_EINVAL, "G * ma too large for the coarse merge");
    QCK(cudaSetDevice(ctx->device));
    QCK(cudaFuncSetAttribute(coarse_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, n_sort * 8));
    for (int q0 = 0; q0 < nq; q0 += kMaxBatch) {   // grid.x limit is not an issue; keep launches bounded like the search
        const int n = std::min(kMaxBatch, nq - q0);
        coarse_merge_kernel<<<n, 256, static_cast<size_t>(n_sort) * 8, ctx->stream>>>(d_keys + static_cast<size_t>(q0) * ma, G, nq,
                                                                                       ma, n_sort,
                                                                                       d_assign + static_cast<size_t>(q0) * ma);
        QCK(cudaGetLastError());
        ctx->launches++;
    }
    return QADC_OK;
}

int qadc_synchronize(qadc_ctx* ctx) {
    if (!ctx) return QADC_EINVAL;
    QCK(cudaSetDevice(ctx->device));
    QCK(cudaStreamSynchronize(ctx->stream));
    if (*ctx->h_err) {
        *ctx->h_err = 0;
        QCK(cudaMemset(ctx->d_err, 0, sizeof(int)));
        return fail(ctx, QADC_EBOUND, "Max quantization bound too high. Try larger keep value.");
    }
    return QADC_OK;
}

int qadc_search(qadc_ctx* ctx, const float* queries, int nq, int ma, int r, uint32_t* out_ids, int8_t* out_dists,
                int32_t* out_counts, qadc_metrics* metrics) {
    int rc = check_search_args(ctx, nq, ma, r);
    if (rc) return rc;
    if (!queries || !out_ids || !out_dists) return fail(ctx, QADC_EINVAL, "null buffer");
    QCK(cudaSetDevice(ctx->device));
    const size_t nr = static_cast<size_t>(nq) * r;
    ENSURE(ctx->b_queries, static_cast<size_t>(nq) * ctx->dim * 4);
    ENSURE(ctx->b_ids, nr * 4); ENSURE(ctx->b_dists, nr); ENSURE(ctx->b_counts, static_cast<size_t>(nq) * 4);
    QCK(cudaEventRecord(ctx->ev[5], ctx->stream));
    QCK(cudaMemcpyAsync(ctx->b_queries.p, queries, static_cast<size_t>(nq) * ctx->dim * 4, cudaMemcpyHostToDevice, ctx->stream));
    rc = qadc_search_device(ctx, ctx->b_queries.as<float>(), nq, ma, r, ctx->b_ids.as<uint32_t>(),
                            ctx->b_dists.as<int8_t>(), ctx->b_counts.as<int32_t>(), nullptr);
    if (rc) return rc;
    QCK(cudaEventRecord(ctx->ev[6], ctx->stream));
    QCK(cudaMemcpyAsync(out_ids, ctx->b_ids.p, nr * 4, cudaMemcpyDeviceToHost, ctx->stream));
    QCK(cudaMemcpyAsync(out_dists, ctx->b_dists.p, nr, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_counts) QCK(cudaMemcpyAsync(out_counts, ctx->b_counts.p, static_cast<size_t>(nq) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    QCK(cudaEventRecord(ctx->ev[7], ctx->stream));
    rc = qadc_synchronize(ctx);
    if (metrics) {
        float ms = 0;
        auto el = [&](int a, int b) { cudaEventElapsedTime(&ms, ctx->ev[a], ctx->ev[b]); return static_cast<double>(ms) * 1e3; };
        metrics->index_us = el(0, 1);
        metrics->rotate_us = 0;   // rotation is fused into the table kernel
        metrics->table_us = el(1, 2);
        metrics->scan_us = el(2, 4);
        metrics->h2d_us = el(5, 0);
        metrics->d2h_us = el(6, 7);
    }
    return rc;
}

int qadc_last_launch_count(const qadc_ctx* ctx) { return ctx ? ctx->launches : 0; }

int qadc_last_scan_ms(qadc_ctx* ctx, float* ms) { return qadc_scan_ms_history(ctx, ms, 1) == 1 ? QADC_OK : QADC_ESTATE; }

int qadc_scan_ms_history(qadc_ctx* ctx, float* ms, int n) {
    if (!ctx || !ms || n <= 0) return QADC_EINVAL;
    if (!ctx->scan_timed) return fail(ctx, QADC_ESTATE, "enable with qadc_set_option(ctx, \"time_scan\", 1)");
    QCK(cudaStreamSynchronize(ctx->stream));
    const long have = std::min<long>(std::min<long>(n, ctx->scan_seq), qadc_ctx::kScanRing);
    for (long i = 0; i < have; ++i) {
        const long seq = ctx->scan_seq - have + i;
        QCK(cudaEventElapsedTime(ms + i, ctx->ev_scan0[seq % qadc_ctx::kScanRing], ctx->ev_scan1[seq % qadc_ctx::kScanRing]));
    }
    return static_cast<int>(have);
}

int qadc_merge_shards_device(qadc_ctx* ctx, const uint64_t* d_keys, const uint32_t* d_ids, int G, int nq, int r,
                             uint32_t* d_out_ids, int8_t* d_out_dists, int32_t* d_out_counts, uint64_t* d_out_keys) {
    if (!ctx || !d_keys || G <= 0 || nq <= 0 || r <= 0 || r > kMergeCap / 2) return fail(ctx, QADC_EINVAL, "bad merge arguments");
    QCK(cudaSetDevice(ctx->device));
    MergeArgs mg{};
    mg.in_keys = d_keys; mg.in_ids = d_ids; mg.L = G; mg.r = r; mg.nq = nq; mg.shard_major = 1;
    mg.out_keys = d_out_keys; mg.out_ids = d_out_ids; mg.out_dists = d_out_dists; mg.out_counts = d_out_counts;
    return run_merge(ctx, mg);
}

int qadc_build_tables(qadc_ctx* ctx, const float* queries, int nq, int ma, int r, const int32_t* assign_in,
                      int32_t* out_assign, float* out_tables, float* out_qmin, float* out_qmax, int8_t* out_qtables) {
    int rc = check_search_args(ctx, nq, ma, r);
    if (rc) return rc;
    if (nq > kMaxBatch) return fail(ctx, QADC_EINVAL, "parity entry points take at most 32768 queries per call");
    if (!queries) return fail(ctx, QADC_EINVAL, "queries is null");
    QCK(cudaSetDevice(ctx->device));
    const size_t nqa = static_cast<size_t>(nq) * ma, td = static_cast<size_t>(ctx->m) * 16;
    ENSURE(ctx->b_queries, static_cast<size_t>(nq) * ctx->dim * 4);
    QCK(cudaMemcpyAsync(ctx->b_queries.p, queries, static_cast<size_t>(nq) * ctx->dim * 4, cudaMemcpyHostToDevice, ctx->stream));
    const int32_t* d_ain = nullptr;
    if (assign_in) {
        ENSURE(ctx->b_ids, nqa * 4);
        QCK(cudaMemcpyAsync(ctx->b_ids.p, assign_in, nqa * 4, cudaMemcpyHostToDevice, ctx->stream));
        d_ain = ctx->b_ids.as<int32_t>();
    }
    ctx->launches = 0;
    rc = tables_device(ctx, ctx->b_queries.as<float>(), nq, ma, r, d_ain, false, out_tables != nullptr);
    if (rc) return rc;
    ctx->sbound_seeded = false; ctx->ghist_zeroed = false;   // no scan follows this call
    QCK(cudaMemcpyAsync(ctx->h_err, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (out_assign) QCK(cudaMemcpyAsync(out_assign, ctx->b_assign.p, nqa * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_tables) QCK(cudaMemcpyAsync(out_tables, ctx->b_tables.p, nqa * td * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_qmin) QCK(cudaMemcpyAsync(out_qmin, ctx->b_qmin.p, static_cast<size_t>(nq) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_qmax) QCK(cudaMemcpyAsync(out_qmax, ctx->b_qmax.p, static_cast<size_t>(nq) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_qtables) QCK(cudaMemcpyAsync(out_qtables, ctx->b_qtables.p, nqa * td, cudaMemcpyDeviceToHost, ctx->stream));
    return qadc_synchronize(ctx);
}

int qadc_scan_with_tables(qadc_ctx* ctx, const int32_t* assign, const int8_t* qtables, int nq, int ma, int r,
                          uint32_t* out_ids, int8_t* out_dists, int32_t* out_counts) {
    int rc = check_search_args(ctx, nq, ma, r);
    if (rc) return rc;
    if (nq > kMaxBatch) return fail(ctx, QADC_EINVAL, "parity entry points take at most 32768 queries per call");
    if (!assign || !qtables || !out_ids || !out_dists) return fail(ctx, QADC_EINVAL, "null buffer");
    const size_t nqa = static_cast<size_t>(nq) * ma, td = static_cast<size_t>(ctx->m) * 16, nr = static_cast<size_t>(nq) * r;
    for (size_t i = 0; i < nqa * td; ++i)
        if (qtables[i] < 0) return fail(ctx, QADC_EINVAL, "int8 table entries must be in [0,127] (db_query_4.cpp:44-55)");
    for (size_t i = 0; i < nqa; ++i)
        if (assign[i] < 0 || assign[i] >= ctx->parts) return fail(ctx, QADC_EINVAL, "assign entry out of range");
    QCK(cudaSetDevice(ctx->device));
    ENSURE(ctx->b_assign, nqa * 4); ENSURE(ctx->b_qtables, nqa * td);
    ENSURE(ctx->b_ids, nr * 4); ENSURE(ctx->b_dists, nr); ENSURE(ctx->b_counts, static_cast<size_t>(nq) * 4);
    QCK(cudaMemcpyAsync(ctx->b_assign.p, assign, nqa * 4, cudaMemcpyHostToDevice, ctx->stream));
    QCK(cudaMemcpyAsync(ctx->b_qtables.p, qtables, nqa * td, cudaMemcpyHostToDevice, ctx->stream));
    ctx->launches = 0;
    rc = scan_device(ctx, ctx->b_assign.as<int32_t>(), ctx->b_qtables.as<int8_t>(), nq, ma, r, ctx->b_ids.as<uint32_t>(),
                     ctx->b_dists.as<int8_t>(), ctx->b_counts.as<int32_t>(), nullptr);
    if (rc) return rc;
    QCK(cudaMemcpyAsync(out_ids, ctx->b_ids.p, nr * 4, cudaMemcpyDeviceToHost, ctx->stream));
    QCK(cudaMemcpyAsync(out_dists, ctx->b_dists.p, nr, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_counts) QCK(cudaMemcpyAsync(out_counts, ctx->b_counts.p, static_cast<size_t>(nq) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    QCK(cudaStreamSynchronize(ctx->stream));
    return QADC_OK;
}

int qadc_dump_distances(qadc_ctx* ctx, int part_i, const int8_t* qtable, int8_t* out) {
    if (!ctx || !ctx->begun) return fail(ctx, QADC_ESTATE, "no database");
    if (part_i < 0 || part_i >= ctx->parts || !qtable || !out) return fail(ctx, QADC_EINVAL, "bad argument");
    const int M = ctx->m;
    for (int i = 0; i < M * 16; ++i)
        if (qtable[i] < 0) return fail(ctx, QADC_EINVAL, "int8 table entries must be in [0,127]");
    const uint32_t size = ctx->h_size[part_i];
    if (size == 0) return QADC_OK;
    QCK(cudaSetDevice(ctx->device));
    ENSURE(ctx->b_dump, static_cast<size_t>(size) + M * 16);
    int8_t* d_t = ctx->b_dump.as<int8_t>();
    int8_t* d_o = d_t + M * 16;
    QCK(cudaMemcpyAsync(d_t, qtable, M * 16, cudaMemcpyHostToDevice, ctx->stream));
    const uint8_t* native = ctx->d_codes + ctx->h_sb_off[part_i] * sb_bytes(M);
    const uint32_t n_sb = (size + kSbVec - 1) / kSbVec;
    if (M == 16) dump_distances_kernel<16><<<(n_sb + 7) / 8, 256, 0, ctx->stream>>>(native, size, d_t, d_o, make_pipek());
    else dump_distances_kernel<32><<<(n_sb + 7) / 8, 256, 0, ctx->stream>>>(native, size, d_t, d_o, make_pipek());
    QCK(cudaGetLastError());
    QCK(cudaMemcpyAsync(out, d_o, size, cudaMemcpyDeviceToHost, ctx->stream));
    QCK(cudaStreamSynchronize(ctx->stream));
    return QADC_OK;
}

int qadc_download_codes(qadc_ctx* ctx, int part_i, uint8_t* out_codes) {
    if (!ctx || !ctx->begun) return fail(ctx, QADC_ESTATE, "no database");
    if (part_i < 0 || part_i >= ctx->parts || !out_codes) return fail(ctx, QADC_EINVAL, "bad argument");
    const int M = ctx->m, CS = M / 2;
    const uint32_t size = ctx->h_size[part_i];
    if (size == 0) return QADC_OK;
    QCK(cudaSetDevice(ctx->device));
    ENSURE(ctx->b_dump, static_cast<size_t>(size) * CS);
    const uint8_t* native = ctx->d_codes + ctx->h_sb_off[part_i] * sb_bytes(M);
    if (M == 16) untranspose_codes_kernel<16><<<(size + 255) / 256, 256, 0, ctx->stream>>>(native, size, ctx->b_dump.as<uint8_t>());
    else untranspose_codes_kernel<32><<<(size + 255) / 256, 256, 0, ctx->stream>>>(native, size, ctx->b_dump.as<uint8_t>());
    QCK(cudaGetLastError());
    QCK(cudaMemcpyAsync(out_codes, ctx->b_dump.p, static_cast<size_t>(size) * CS, cudaMemcpyDeviceToHost, ctx->stream));
    QCK(cudaStreamSynchronize(ctx->stream));
    return QADC_OK;
}

int qadc_encode(qadc_ctx* ctx, const float* vectors, uint32_t count, int32_t* out_assign, uint8_t* out_codes) {
    if (!ctx || !vectors || !out_codes) return fail(ctx, QADC_EINVAL, "null buffer");
    if (ctx->m == 0) return fail(ctx, QADC_ESTATE, "qadc_set_pq must be called first");
    if (count == 0) return QADC_OK;
    QCK(cudaSetDevice(ctx->device));
    const int dim = ctx->dim, M = ctx->m, CS = M * ctx->bits / 8;
    const bool ivf = ctx->K > 0;
    const uint32_t kChunk = 1u << 20;
    ENSURE(ctx->b_queries, static_cast<size_t>(std::min(count, kChunk)) * dim * 4);
    ENSURE(ctx->b_tables, static_cast<size_t>(std::min(count, kChunk)) * dim * 4);   // rotated copy
    ENSURE(ctx->b_assign, static_cast<size_t>(std::min(count, kChunk)) * 4);
    ENSURE(ctx->b_dump, static_cast<size_t>(std::min(count, kChunk)) * CS);
    for (uint32_t off = 0; off < count; off += kChunk) {
        const uint32_t n = std::min(kChunk, count - off);
        float* d_x = ctx->b_queries.as<float>();
        QCK(cudaMemcpyAsync(d_x, vectors + static_cast<size_t>(off) * dim, static_cast<size_t>(n) * dim * 4,
                            cudaMemcpyHostToDevice, ctx->stream));
        int32_t* d_assign = nullptr;
        if (ivf) {
            // index_db::assign_single_compute_residuals (databases.hpp:252-268): nearest cell, k = 1
            d_assign = ctx->b_assign.as<int32_t>();
            const int chunk_q = std::max(1, static_cast<int>(std::min<size_t>(n, (size_t(1) << 28) / ctx->K)));
            ENSURE(ctx->b_cdist, static_cast<size_t>(chunk_q) * ctx->K * 4);
            for (uint32_t q0 = 0; q0 < n; q0 += chunk_q) {
                const int nn = static_cast<int>(std::min<uint32_t>(chunk_q, n - q0));
                dim3 grid((nn + kCoarseTQ - 1) / kCoarseTQ, (ctx->K + kCoarseTC - 1) / kCoarseTC);
                coarse_dist_kernel<<<grid, 256, 0, ctx->stream>>>(d_x + static_cast<size_t>(q0) * dim, nn, dim, ctx->d_centroids,
                                                                  ctx->K, ctx->b_cdist.as<float>());
                coarse_select_kernel<<<nn, kSelThreads, 0, ctx->stream>>>(ctx->b_cdist.as<float>(), ctx->K, 1, d_assign + q0, nullptr, 0u);
                QCK(cudaGetLastError());
            }
            if (out_assign)
                QCK(cudaMemcpyAsync(out_assign + off, d_assign, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost, ctx->stream));
        }
        const float* d_in = d_x;
        bool residual_done = false;
        if (ctx->d_rotation) {
            // OPQ: rotate the vector — or, with inverted lists, its residual (residual -> rotate -> encode)
            const size_t tot = static_cast<size_t>(n) * dim;
            rotate_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, ctx->stream>>>(
                d_x, n, dim, ctx->d_rotation, ivf ? ctx->d_centroids : nullptr, d_assign, ctx->b_tables.as<float>());
            QCK(cudaGetLastError());
            d_in = ctx->b_tables.as<float>();
            residual_done = ivf;
        }
        // one thread per (vector, code byte); 16-bit codes: per (vector, sub-quantiser), two bytes each
        const size_t threads = ctx->bits == 16 ? static_cast<size_t>(n) * M : static_cast<size_t>(n) * CS;
        encode_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, ctx->stream>>>(
            d_in, n, dim, M, ctx->bits, ctx->d_codebooks, (ivf && !residual_done) ? ctx->d_centroids : nullptr, d_assign,
            ctx->b_dump.as<uint8_t>());
        QCK(cudaGetLastError());
        QCK(cudaMemcpyAsync(out_codes + static_cast<size_t>(off) * CS, ctx->b_dump.p, static_cast<size_t>(n) * CS,
                            cudaMemcpyDeviceToHost, ctx->stream));
        QCK(cudaStreamSynchronize(ctx->stream));
    }
    return QADC_OK;
}

// ---- plain ADC (db_query): row-major database + float scan ---------------------------------
int qadc_adc_load(qadc_ctx* ctx, int partition_count, const uint64_t* offsets, const uint8_t* codes,
                  const uint32_t* labels) {
    if (!ctx || !offsets || partition_count <= 0) return fail(ctx, QADC_EINVAL, "bad database description");
    if (ctx->m == 0) return fail(ctx, QADC_ESTATE, "qadc_set_pq must be called first");
    if (ctx->K == 0 && partition_count != 1) return fail(ctx, QADC_EINVAL, "a flat database has exactly one partition");
    if (ctx->K > 0 && partition_count != ctx->K) return fail(ctx, QADC_EINVAL, "partition_count != coarse centroid count");
    if (ctx->K > 0 && !labels) return fail(ctx, QADC_EINVAL, "inverted lists need labels");
    for (int p = 0; p < partition_count; ++p)
        if (offsets[p + 1] < offsets[p]) return fail(ctx, QADC_EINVAL, "offsets must be non-decreasing");
    const uint64_t n = offsets[partition_count] - offsets[0];
    if (n > 0 && !codes) return fail(ctx, QADC_EINVAL, "codes is null");
    if (n >= (uint64_t(1) << 32)) return fail(ctx, QADC_EINVAL, "more than 2^32 - 1 vectors");
    QCK(cudaSetDevice(ctx->device));
    cudaFree(ctx->d_rows); cudaFree(ctx->d_row_off); cudaFree(ctx->d_adc_labels);
    ctx->d_rows = nullptr; ctx->d_row_off = nullptr; ctx->d_adc_labels = nullptr; ctx->adc_parts = 0;
    const size_t cs = static_cast<size_t>(ctx->m) * ctx->bits / 8;
    std::vector<uint64_t> rel(partition_count + 1);
    for (int p = 0; p <= partition_count; ++p) rel[p] = offsets[p] - offsets[0];
    QCK(cudaMalloc(&ctx->d_rows, std::max<size_t>(16, n * cs)));
    QCK(cudaMalloc(&ctx->d_row_off, rel.size() * 8));
    QCK(cudaMemcpy(ctx->d_row_off, rel.data(), rel.size() * 8, cudaMemcpyHostToDevice));
    if (n) QCK(cudaMemcpy(ctx->d_rows, codes + offsets[0] * cs, n * cs, cudaMemcpyHostToDevice));
    if (labels) {
        QCK(cudaMalloc(&ctx->d_adc_labels, std::max<size_t>(4, n * 4)));
        if (n) QCK(cudaMemcpy(ctx->d_adc_labels, labels + offsets[0], n * 4, cudaMemcpyHostToDevice));
    }
    ctx->adc_parts = partition_count;
    return QADC_OK;
}

int qadc_adc_search(qadc_ctx* ctx, const float* queries, int nq, int ma, int r, uint32_t* out_ids, float* out_dists,
                    int32_t* out_counts) {
    if (!ctx) return QADC_EINVAL;
    if (ctx->adc_parts == 0) return fail(ctx, QADC_ESTATE, "qadc_adc_load must be called first");
    if (!queries || !out_ids || !out_dists) return fail(ctx, QADC_EINVAL, "null buffer");
    if (nq <= 0 || ma <= 0 || r <= 0) return fail(ctx, QADC_EINVAL, "nq, ma and r must be positive");
    if (r > kSelCap / 2) return fail(ctx, QADC_EINVAL, "r > 1024 is not supported");
    if (ctx->K == 0 && ma != 1) return fail(ctx, QADC_EINVAL, "a flat database must be queried with ma = 1 (SURVEY App. B)");
    if (ctx->K > 0 && (ma > ctx->K || ma > kSelCap / 2)) return fail(ctx, QADC_EINVAL, "ma exceeds the partition count (or 1024)");
    QCK(cudaSetDevice(ctx->device));
    const int M = ctx->m, dim = ctx->dim, bits = ctx->bits;
    const size_t td = static_cast<size_t>(M) << bits;   // floats per table
    const bool flat = ctx->K == 0;
    // query sub-batches: float tables <= 1 GiB, grids below 65 536
    const int bq = static_cast<int>(std::max<size_t>(1, std::min<size_t>(std::min(nq, kMaxBatch), (size_t(1) << 28) / (td * ma))));
    const int nsplit = std::max(1, std::min(64, (2 * ctx->sm_count + bq - 1) / bq));
    ctx->launches = 0;
    for (int q0 = 0; q0 < nq; q0 += bq) {
        const int n = std::min(bq, nq - q0);
        const size_t nqa = static_cast<size_t>(n) * ma;
        ENSURE(ctx->b_queries, static_cast<size_t>(n) * dim * 4);
        ENSURE(ctx->b_assign, nqa * 4);
        ENSURE(ctx->b_tables, nqa * td * 4);
        ENSURE(ctx->b_tmin, nqa * 4);
        ENSURE(ctx->b_plists, static_cast<size_t>(n) * nsplit * r * 8);
        ENSURE(ctx->b_keys, static_cast<size_t>(n) * r * 8);
        ENSURE(ctx->b_ids, static_cast<size_t>(n) * r * 4);
        ENSURE(ctx->b_adc_dists, static_cast<size_t>(n) * r * 4);
        ENSURE(ctx->b_counts, static_cast<size_t>(n) * 4);
        float* d_q = ctx->b_queries.as<float>();
        int32_t* d_assign = ctx->b_assign.as<int32_t>();
        QCK(cudaMemcpyAsync(d_q, queries + static_cast<size_t>(q0) * dim, static_cast<size_t>(n) * dim * 4,
                            cudaMemcpyHostToDevice, ctx->stream));
        if (flat) {
            QCK(cudaMemsetAsync(d_assign, 0, nqa * 4, ctx->stream));
        } else {
            const int chunk_q = std::max(1, static_cast<int>(std::min<size_t>(n, (size_t(1) << 28) / ctx->K)));
            ENSURE(ctx->b_cdist, static_cast<size_t>(chunk_q) * ctx->K * 4);
            for (int c0 = 0; c0 < n; c0 += chunk_q) {
                const int cn = std::min(chunk_q, n - c0);
                dim3 grid((cn + kCoarseTQ - 1) / kCoarseTQ, (ctx->K + kCoarseTC - 1) / kCoarseTC);
                coarse_dist_kernel<<<grid, 256, 0, ctx->stream>>>(d_q + static_cast<size_t>(c0) * dim, cn, dim, ctx->d_centroids,
                                                                  ctx->K, ctx->b_cdist.as<float>());
                QCK(cudaGetLastError());
                coarse_select_kernel<<<cn, kSelThreads, 0, ctx->stream>>>(ctx->b_cdist.as<float>(), ctx->K, ma,
                                                                         d_assign + static_cast<size_t>(c0) * ma, nullptr, 0u);
                QCK(cudaGetLastError());
                ctx->launches += 2;
            }
        }
        {
            dim3 tgrid((ma + 7) / 8, n);
            auto launch_tables = [&](auto kernel) {
                if (static_cast<size_t>(dim) * 64 > 48 * 1024)
                    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dim * 64);
                kernel<<<tgrid, 256, static_cast<size_t>(dim) * 8 * 8, ctx->stream>>>(
                    d_q, dim, M, ctx->d_codebooks, ctx->d_rotation, flat ? nullptr : ctx->d_centroids, d_assign, ma,
                    ctx->b_tables.as<float>(), ctx->b_tmin.as<float>(), bits, nullptr);
            };
            switch (dim / M) {
                case 2: launch_tables(tables_kernel<2>); break;
                case 3: launch_tables(tables_kernel<3>); break;
                case 4: launch_tables(tables_kernel<4>); break;
                case 6: launch_tables(tables_kernel<6>); break;
                case 8: launch_tables(tables_kernel<8>); break;
                case 12: launch_tables(tables_kernel<12>); break;
                case 16: launch_tables(tables_kernel<16>); break;
                default: launch_tables(tables_kernel<0>); break;
            }
            ctx->launches++;
            QCK(cudaGetLastError());
        }
        AdcScanArgs a;
        a.rows = ctx->d_rows; a.row_off = ctx->d_row_off; a.assign = d_assign; a.tables = ctx->b_tables.as<float>();
        a.ma = ma; a.r = r; a.nsplit = nsplit; a.lists = ctx->b_plists.as<uint64_t>();
        dim3 sgrid(nsplit, n);
        if (bits == 4 && M == 16) adc_scan_kernel<4, 16><<<sgrid, kSelThreads, 0, ctx->stream>>>(a);
        else if (bits == 4 && M == 32) adc_scan_kernel<4, 32><<<sgrid, kSelThreads, 0, ctx->stream>>>(a);
        else if (bits == 8 && M == 4) adc_scan_kernel<8, 4><<<sgrid, kSelThreads, 0, ctx->stream>>>(a);
        else if (bits == 8 && M == 8) adc_scan_kernel<8, 8><<<sgrid, kSelThreads, 0, ctx->stream>>>(a);
        else if (bits == 8) adc_scan_kernel<8, 16><<<sgrid, kSelThreads, 0, ctx->stream>>>(a);
        else if (M == 2) adc_scan_kernel<16, 2><<<sgrid, kSelThreads, 0, ctx->stream>>>(a);
        else if (M == 4) adc_scan_kernel<16, 4><<<sgrid, kSelThreads, 0, ctx->stream>>>(a);
        else adc_scan_kernel<16, 8><<<sgrid, kSelThreads, 0, ctx->stream>>>(a);
        ctx->launches++;
        QCK(cudaGetLastError());
        MergeArgs mg{};
        mg.in_keys = ctx->b_plists.as<uint64_t>(); mg.L = nsplit; mg.r = r; mg.nq = n;
        mg.out_keys = ctx->b_keys.as<uint64_t>(); mg.out_counts = ctx->b_counts.as<int32_t>();
        int rc = run_merge(ctx, mg);
        if (rc) return rc;
        adc_finalize_kernel<<<n, 256, 0, ctx->stream>>>(ctx->b_keys.as<uint64_t>(), r, ma, d_assign, ctx->d_row_off,
                                                        ctx->d_adc_labels, ctx->b_ids.as<uint32_t>(),
                                                        ctx->b_adc_dists.as<float>());
        ctx->launches++;
        QCK(cudaGetLastError());
        QCK(cudaMemcpyAsync(out_ids + static_cast<size_t>(q0) * r, ctx->b_ids.p, static_cast<size_t>(n) * r * 4,
                            cudaMemcpyDeviceToHost, ctx->stream));
        QCK(cudaMemcpyAsync(out_dists + static_cast<size_t>(q0) * r, ctx->b_adc_dists.p, static_cast<size_t>(n) * r * 4,
                            cudaMemcpyDeviceToHost, ctx->stream));
        if (out_counts)
            QCK(cudaMemcpyAsync(out_counts + q0, ctx->b_counts.p, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost, ctx->stream));
        QCK(cudaStreamSynchronize(ctx->stream));   // scratch is reused by the next sub-batch
    }
    return QADC_OK;
}

int qadc_set_option(qadc_ctx* ctx, const char* key, long value) {
    if (!ctx || !key) return QADC_EINVAL;
    if (!strcmp(key, "flat_qb")) ctx->opt_flat_qb = value;
    else if (!strcmp(key, "flat_chunks")) ctx->opt_flat_chunks = value;
    else if (!strcmp(key, "flat_filter")) ctx->opt_flat_filter = value != 0;
    else if (!strcmp(key, "ivf_fused")) ctx->opt_ivf_fused = value;
    else if (!strcmp(key, "flat_ring")) ctx->opt_flat_ring = value != 0;
    else if (!strcmp(key, "flat_seed")) ctx->opt_flat_seed = value != 0;
    else if (!strcmp(key, "flat_prep")) ctx->opt_flat_prep = value != 0;
    else if (!strcmp(key, "flat_share")) ctx->opt_flat_share = value != 0;
#ifdef QADC_EXPERIMENT
    else if (!strcmp(key, "seed_minus")) ctx->opt_seed_minus = value;   // tools/exp_seed.py only: a seed below the true bound loses results
#endif
    else if (!strcmp(key, "ivf_sb_per_item")) {
        if (value < 1 || value > (1 << 20)) return fail(ctx, QADC_EINVAL, "ivf_sb_per_item out of range");
        ctx->ivf_sb_per_item = static_cast<int>(value);
    }
    else if (!strcmp(key, "time_scan")) ctx->scan_timed = value != 0;
    else return fail(ctx, QADC_EINVAL, std::string("unknown option ") + key);
    return QADC_OK;
}

}  // extern "C"

#include "qadc_multi.cuh"
