// qadc_flatprep.cuh — qmax of a FLAT database with a long keep-prefix: Quick ADC applied to its own prefix.
//
// scanner_4::query_scan (db_query_4.cpp:245-284) takes qmax = the r-th smallest FLOAT ADC distance among the prefix
// vectors (query_scan_start -> scan_4, query_common.hpp:59-90).  Evaluated literally that is 16 dependent table
// look-ups per (prefix vector, query): 500 000 x 16 queries cost 53 us per step of the 1e9 bench (137 instructions per
// pair), another 31 us go into the int8 histogram of the same vectors that seeds the scan's bound, and neither shrinks
// when the database is sharded — every shard scans the whole replicated prefix.  Here only a sample and a few thousand
// candidates per query are evaluated in float; the rest of the prefix is excluded by an int8 LOWER bound of its
// distances, computed with the PRMT core of the main scan on a nibble-plane copy of the prefix:
//
//   1. flat_prefix_sample_kernel (grid = (sample / 2048, queries)): exact distances of the first 65 536 prefix vectors,
//      the r smallest of every 2048.
//   2. flat_prefix_bound_kernel: U0 = r-th smallest of those (an upper bound of qmax: the sample alone holds r vectors
//      at or below it) and provisional int8 tables  P[j][c] = floor((T[j][c] - min_j) * 127 / (U0 - sum_j min_j) - 0.05)
//      clamped to [0, 127].  With L(v) = sum_j P[j][c_j(v)]:  dist(v) <= U0  =>  L(v) <= 127  (every term is rounded down
//      by at least 0.05 of a step, float rounding moves a term by less than 0.02 of a step when the range U0 - sum min
//      is at least a hundredth of U0; smaller ranges fall back to all-zero tables, for which every vector passes).
//   3. flat_prefix_pass_kernel<M, 1>: the positions with L <= 127 — a third of a percent of the prefix on the bench —
//      are the candidates.
//   4. flat_bounds_final_kernel: exact float distances of the candidates in sub-quantiser order (the same __fadd_rn
//      chain as prefix_scan_kernel and oracle/qadc_oracle.c), r-th smallest = qmax — the candidates contain every
//      vector at or below U0 >= qmax, so this IS the r-th smallest of the whole prefix, bit for bit — then QuantizerMAX
//      (quantize_query), then the scan's shared-bound seed: the largest int8 distance among the candidates at or below
//      qmax (at least r scanned vectors are at or below it; the histogram-based seed of prefix_hist_kernel is 1-3 lower
//      on the bench, at the price of another pass over the prefix).  More candidates than slots (degenerate data):
//      every prefix vector is evaluated, slow but exact.
//
// flat_prefix_pass_kernel<M, 0> (histogram of the int8 sums + its r-th smallest by the last CTA to finish) seeds the
// bound when the int8 tables are injected (qadc_scan_with_tables) and was the first version's threshold pass.
#pragma once
#include "qadc_scan.cuh"
#include "qadc_tables.cuh"

namespace qadc {

constexpr int kPrefSplit = 2048;              // sample vectors per CTA of the sample kernel
constexpr uint32_t kPrefSampleMax = 65536;    // sample = the first min(n_prefix, 65 536) prefix vectors, whole splits
constexpr uint32_t kPrefMinNative = 131072;   // shorter prefixes keep the plain float scan
constexpr int kPrefMaxR = 512;                // r <= a quarter of a split
constexpr int kPrefCandCap = 16384;           // candidate slots per query (fewer for large batches: FlatPrepArgs::cand_cap)
constexpr int kPrefFastCap = 2 * kSelCap;     // candidates whose distances fit one shared-memory buffer

struct FlatPrepArgs {
    const uint8_t* starts;        // row-major prefix codes
    const uint8_t* native;        // the same vectors as nibble-plane superblocks
    uint32_t n_prefix;
    const float* tables;          // [nq][M*16]
    int r, nsplit;
    uint32_t* sample_lists;       // [nq][nsplit][r] float bits: the r smallest distances of every sample split
    int8_t* prov_qt;              // [nq][M*16] provisional lower-bound tables
    unsigned int* cand_count;     // [nq] (zeroed by the caller)
    uint32_t* cand;               // [nq][cand_cap] prefix positions
    uint32_t cand_cap;            // candidate slots per query
    int* seed_out;                // [nq] the scan's shared bound
};

// exact float ADC distance of one row-major code: sub-quantiser order, __fadd_rn (query_common.hpp:72-80)
template <int M>
__device__ __forceinline__ float prefix_float_distance(const uint32_t (&w)[M / 8], const float* tab) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < M; ++j) s = __fadd_rn(s, tab[j * 16 + ((w[j >> 3] >> (4 * (j & 7))) & 15u)]);
    return s;
}
template <int M>
__device__ __forceinline__ int prefix_int8_distance(const uint32_t (&w)[M / 8], const int8_t* tab) {
    int s = 0;
#pragma unroll
    for (int j = 0; j < M; ++j) s += tab[j * 16 + ((w[j >> 3] >> (4 * (j & 7))) & 15u)];
    return s;
}
template <int M>
__device__ __forceinline__ void load_row_code(const uint8_t* codes, size_t v, uint32_t (&w)[M / 8]) {
    if constexpr (M == 16) {
        const uint2 c = *reinterpret_cast<const uint2*>(codes + v * 8);
        w[0] = c.x; w[1] = c.y;
    } else {
        const uint4 c = *reinterpret_cast<const uint4*>(codes + v * 16);
        w[0] = c.x; w[1] = c.y; w[2] = c.z; w[3] = c.w;
    }
}

// grid = (sample splits, queries): the r smallest exact distances of prefix vectors [split * 2048, (split + 1) * 2048)
template <int M>
__global__ void __launch_bounds__(kSelThreads) flat_prefix_sample_kernel(const FlatPrepArgs a) {
    constexpr int TE = M * 16, VPT = kPrefSplit / kSelThreads;
    __shared__ float tab[TE];
    __shared__ uint32_t vals[kPrefSplit];
    __shared__ int hist[256], state[2], count;
    const int split = blockIdx.x, q = blockIdx.y, tid = threadIdx.x;
    for (int i = tid; i < TE; i += kSelThreads) tab[i] = a.tables[static_cast<size_t>(q) * TE + i];
    uint32_t w[VPT][M / 8];
#pragma unroll
    for (int u = 0; u < VPT; ++u)   // (the host sizes the sample in whole splits inside the prefix)
        load_row_code<M>(a.starts, static_cast<size_t>(split) * kPrefSplit + tid + u * kSelThreads, w[u]);
    if (tid == 0) count = 0;
    __syncthreads();
#pragma unroll
    for (int u = 0; u < VPT; ++u) vals[tid + u * kSelThreads] = __float_as_uint(prefix_float_distance<M>(w[u], tab));
    __syncthreads();
    int n_less;
    const uint32_t b = block_radix_select(vals, kPrefSplit, a.r, hist, state, tid, n_less);
    uint32_t* dst = a.sample_lists + (static_cast<size_t>(q) * a.nsplit + split) * a.r;
    for (int i = tid; i < kPrefSplit; i += kSelThreads) {
        const uint32_t v = vals[i];
        if (v < b) dst[atomicAdd(&count, 1)] = v;   // n_less values, in any order
    }
    for (int i = n_less + tid; i < a.r; i += kSelThreads) dst[i] = b;   // the r-th value and its ties
}

// grid = queries: U0 and the provisional lower-bound tables
template <int M>
__global__ void __launch_bounds__(kSelThreads) flat_prefix_bound_kernel(const FlatPrepArgs a) {
    constexpr int TE = M * 16;
    __shared__ float tab[TE];
    __shared__ int hist[256], state[2];
    __shared__ float mn[M];
    const int q = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < TE; i += kSelThreads) tab[i] = a.tables[static_cast<size_t>(q) * TE + i];
    __syncthreads();
    if (tid < M) {
        float m = tab[tid * 16];
#pragma unroll
        for (int c = 1; c < 16; ++c) m = fminf(m, tab[tid * 16 + c]);
        mn[tid] = m;
    }
    int n_less;
    const float u0 = __uint_as_float(block_radix_select(a.sample_lists + static_cast<size_t>(q) * a.nsplit * a.r, a.nsplit * a.r,
                                                        a.r, hist, state, tid, n_less));   // (ends with a block barrier: mn is visible)
    float smin = 0.f;
#pragma unroll
    for (int j = 0; j < M; ++j) smin += mn[j];
    const float range = u0 - smin;
    // the bound needs a usable scale: a positive range that is not lost in the rounding of u0 and smin
    // (a hundredth of u0: the rounding of the 16-term sums is then below 0.02 of a table step, within the 0.05 slack)
    const bool ok = u0 < 1e30f && range > 0.f && range >= 1e-2f * u0;
    const float scale = ok ? 127.0f / range : 0.f;
    for (int e = tid; e < TE; e += kSelThreads) {
        int qv = 0;
        if (ok) {
            const float x = (tab[e] - mn[e >> 4]) * scale - 0.05f;
            qv = x <= 0.f ? 0 : (x >= 127.f ? 127 : static_cast<int>(x));
        }
        a.prov_qt[static_cast<size_t>(q) * TE + e] = static_cast<int8_t>(qv);
    }
}

// grid = (splits, queries), 8 warps; a warp takes whole superblocks.
// MODE 0: hist[q][s] += 1 for every prefix vector whose int8 sum s is at most 127; the LAST CTA of a query to finish
//         (ticket counter done[q]) scans the histogram and stores rth[q] = min(rth_cap, the smallest s whose cumulative
//         count reaches r, 127 when none does) — the r-th smallest sum, without a launch of its own.  With the final int8
//         tables and rth_cap = 126 that is the scan's shared-bound seed (what prefix_hist_kernel + prefix_bound_kernel
//         compute).
// MODE 1: the positions whose sum is at most 127 (or min(127, rth[q] + margin) when rth is given) are appended to the
//         query's candidate list.
template <int M, int MODE>
__global__ void __launch_bounds__(256) flat_prefix_pass_kernel(const uint8_t* __restrict__ native, uint32_t n_prefix,
                                                               const int8_t* __restrict__ qt, unsigned int* hist,
                                                               unsigned int* done, int* rth, int rth_cap, int r, int margin,
                                                               unsigned int* __restrict__ cand_count,
                                                               uint32_t* __restrict__ cand, uint32_t cand_cap, const PipeK pk) {
    constexpr int kQuads = M / 4, kSbBytes = M * 128;
    __shared__ uint4 tab[M];
    __shared__ unsigned int h[128];
    __shared__ int s_last;
    const int q = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < M) tab[tid] = reinterpret_cast<const uint4*>(qt)[static_cast<size_t>(q) * M + tid];
    if (MODE == 0 && tid < 128) h[tid] = 0;
    __syncthreads();
    const uint32_t thr = (MODE == 0 || !rth) ? 127u : static_cast<uint32_t>(min(127, rth[q] + margin));
    const uint32_t n_sb = (n_prefix + kSbVec - 1) / kSbVec;
    for (uint32_t sb = blockIdx.x * 8 + warp; sb < n_sb; sb += gridDim.x * 8) {
        const uint4* src = reinterpret_cast<const uint4*>(native + static_cast<size_t>(sb) * kSbBytes) + lane;
        uint4 w[kQuads];
#pragma unroll
        for (int qd = 0; qd < kQuads; ++qd) w[qd] = __ldg(src + qd * 32);
        GroupAcc g;
        acc_init(g, 0);
#pragma unroll
        for (int qd = 0; qd < kQuads; ++qd) {
            const uint4 tq[4] = {tab[4 * qd], tab[4 * qd + 1], tab[4 * qd + 2], tab[4 * qd + 3]};
            lut_quad(w[qd], tq, g, pk);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t pos = sb * kSbVec + lane * 8 + k;
            const uint32_t s = lane_sum(g, k, 0);
            if (pos < n_prefix && s <= thr) {
                if (MODE == 0) {
                    atomicAdd(&h[s], 1u);
                } else {
                    const unsigned int slot = atomicAdd(&cand_count[q], 1u);
                    if (slot < cand_cap) cand[static_cast<size_t>(q) * cand_cap + slot] = pos;
                }
            }
        }
    }
    if (MODE == 0) {
        __syncthreads();
        unsigned int* gh = hist + static_cast<size_t>(q) * 128;
        if (tid < 128 && h[tid]) atomicAdd(&gh[tid], h[tid]);
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = atomicAdd(&done[q], 1u) == gridDim.x - 1;
        __syncthreads();
        if (s_last && tid < 32) {   // warp 0 of the query's last CTA: every CTA's counts are in (fence + ticket)
            __threadfence();
            unsigned int c[4], mine = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) { c[i] = __ldcg(&gh[tid * 4 + i]); mine += c[i]; }
            unsigned int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o);
                if (tid >= o) incl += up;
            }
            const unsigned reached = __ballot_sync(0xffffffffu, incl >= static_cast<unsigned int>(r));
            if (reached == 0) {
                if (tid == 0) rth[q] = min(rth_cap, 127);
            } else if (tid == __ffs(reached) - 1) {
                unsigned int cum = incl - mine;
                int b = 0;
                while (cum + c[b] < static_cast<unsigned int>(r)) { cum += c[b]; ++b; }
                rth[q] = min(rth_cap, tid * 4 + b);
            }
        }
    }
}

// grid = queries: exact distances of the candidates -> qmax, QuantizerMAX for the query's table, the scan's bound seed
template <int M>
__global__ void __launch_bounds__(kSelThreads) flat_bounds_final_kernel(const FlatPrepArgs a, float* tables_rw,
                                                                       const float* __restrict__ tmin,
                                                                       int8_t* qtables, float* __restrict__ qmin_out,
                                                                       float* __restrict__ qmax_out, int* __restrict__ err) {
    constexpr int TE = M * 16, VPT = kSelCap / 2 / kSelThreads;
    __shared__ uint32_t vbuf[2][kSelCap];   // fast path: one buffer of kPrefFastCap distances, aligned with the candidate list
    __shared__ float tab[TE];
    __shared__ __align__(16) int8_t qtab[TE];
    __shared__ int count, hist[256], state[2];
    __shared__ unsigned int bound;
    __shared__ float red[8];
    __shared__ int s_seed;
    const int q = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < TE; i += kSelThreads) tab[i] = a.tables[static_cast<size_t>(q) * TE + i];
    if (tid == 0) s_seed = 0;
    const unsigned int n_found = a.cand_count[q];
    const bool all = n_found > a.cand_cap;   // more candidates than slots: evaluate every prefix vector
    const uint32_t n = all ? a.n_prefix : n_found;
    const uint32_t* cand = a.cand + static_cast<size_t>(q) * a.cand_cap;
    const bool fast = !all && n <= static_cast<uint32_t>(kPrefFastCap);
    uint32_t* vals = &vbuf[0][0];
    float qmax;
    if (fast) {
        // the usual case (a few thousand candidates): all distances into one buffer, one selection
        __syncthreads();   // tab
        for (uint32_t base = 0; base < n; base += 4 * kSelThreads) {
            uint32_t pos[4], w[4][M / 8];
#pragma unroll
            for (int u = 0; u < 4; ++u) pos[u] = (base + tid + u * kSelThreads < n) ? cand[base + tid + u * kSelThreads] : 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u) load_row_code<M>(a.starts, pos[u], w[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (base + tid + u * kSelThreads < n)
                    vals[base + tid + u * kSelThreads] = __float_as_uint(prefix_float_distance<M>(w[u], tab));
        }
        __syncthreads();
        qmax = 3.402823466e+38f;
        if (n >= static_cast<uint32_t>(a.r)) {   // block-uniform
            int n_less;
            qmax = __uint_as_float(block_radix_select(vals, static_cast<int>(n), a.r, hist, state, tid, n_less));
        }
    } else {
        BlockMinValues top{vbuf[0], vbuf[1], &count, hist, state, &bound};
        top.init(tid);   // (a block barrier: tab is visible)
        for (uint32_t base = 0; base < n; base += kSelCap / 2) {
            const uint32_t end = min(base + kSelCap / 2, n);
            uint32_t w[VPT][M / 8];
#pragma unroll
            for (int u = 0; u < VPT; ++u) {
                const uint32_t i = base + tid + u * kSelThreads;
                if (i < end) load_row_code<M>(a.starts, all ? i : cand[i], w[u]);
            }
#pragma unroll
            for (int u = 0; u < VPT; ++u)
                if (base + tid + u * kSelThreads < end) top.push(__float_as_uint(prefix_float_distance<M>(w[u], tab)));
            top.maybe_compact(a.r, tid, false);
        }
        top.maybe_compact(a.r, tid, true);
        qmax = (count >= a.r) ? __uint_as_float(bound) : 3.402823466e+38f;
    }
    if (tid == 0) qmax_out[q] = qmax;
    quantize_query(tables_rw, tmin, qmax, 1, M, qtables, qmin_out, err, nullptr, nullptr, nullptr, q, tid, red);
    // seed of the scan's shared bound: the candidates at or below qmax are at least r scanned vectors, so nothing farther
    // than the largest of THEIR int8 distances can reach the top r (126 = "anything below 127" when that is not known)
    int seed = 126;
    if (fast && n >= static_cast<uint32_t>(a.r)) {
        __syncthreads();   // this CTA's int8 table is in global memory
        for (int i = tid; i < TE / 16; i += kSelThreads)
            reinterpret_cast<uint4*>(qtab)[i] = reinterpret_cast<const uint4*>(qtables + static_cast<size_t>(q) * TE)[i];
        __syncthreads();
        const uint32_t qbits = __float_as_uint(qmax);
        int mx = 0;
        for (uint32_t i = tid; i < n; i += kSelThreads) {
            if (vals[i] <= qbits) {
                uint32_t w[M / 8];
                load_row_code<M>(a.starts, cand[i], w);
                mx = max(mx, prefix_int8_distance<M>(w, qtab));
            }
        }
        if (mx > 0) atomicMax(&s_seed, mx);
        __syncthreads();
        seed = min(126, s_seed);
    }
    if (tid == 0) a.seed_out[q] = seed;
}

}  // namespace qadc
