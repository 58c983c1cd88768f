// qadc_tables.cuh — per-query table pipeline on the device (sm_100a):
//   coarse assignment      find_k_neighbors (neighbors.cpp:30-76, with the :64 stride fixed)
//   residuals              substract_vectors_from_unique (databases.cpp:37-48)
//   OPQ rotation           opq::rotate_multiple_vectors (quantizers.hpp:289-301)
//   float lookup tables    compute_dists_single_simd_cg / fmanorm (distances.hpp:60-77, :294-311)
//   keep-prefix float ADC  scan_4 + scanner_4::query_scan_start (query_common.hpp:59-90,
//                          db_query_4.cpp:230-242) -> qmax = r-th smallest prefix distance
//   bounds + int8 tables   scanner_4::query_scan (db_query_4.cpp:256-284), QuantizerMAX (:37-71)
// Every float operation is an explicit round-to-nearest intrinsic in a fixed order, the same
// order oracle/qadc_oracle.c uses, so these stages are bit-reproducible against the oracle
// (and within ~1e-6 relative of the -ffast-math reference build).
#pragma once
#include "qadc_device.cuh"

namespace qadc {

constexpr int kSelCap = 2048;       // streaming top-k buffer of the selection kernels
constexpr int kSelThreads = 256;

// Block-wide streaming "keep the k smallest u64 keys" helper over a shared buffer.
struct BlockTopK {
    uint64_t* keys;                  // kSelCap
    int* count;
    unsigned long long* bound_key;
    __device__ __forceinline__ void init(int tid) {
        for (int i = tid; i < kSelCap; i += kSelThreads) keys[i] = kEmptyKey;
        if (tid == 0) { *count = 0; *bound_key = kEmptyKey; }
        __syncthreads();
    }
    __device__ __forceinline__ void push(uint64_t k) {
        if (k < *bound_key) keys[atomicAdd(count, 1)] = k;
    }
    // call after every round of <= kSelCap/2 pushes (all threads); force = last round
    __device__ __forceinline__ void maybe_compact(int k, int tid, bool force) {
        __syncthreads();
        const int c = *count;
        __syncthreads();   // every thread has read the count before anyone pushes again (uniform decision)
        if (c > kSelCap / 2 || force) {
            int n_sort = 64;   // only the occupied power-of-two prefix needs sorting (the rest is empty)
            while (n_sort < c) n_sort <<= 1;
            bitonic_sort_u64(keys, n_sort, tid, kSelThreads, BlockSync());
            const int n = min(c, k);
            for (int i = k + tid; i < n_sort; i += kSelThreads) keys[i] = kEmptyKey;
            if (tid == 0) { *count = n; *bound_key = (n == k) ? keys[k - 1] : kEmptyKey; }
            __syncthreads();
        }
    }
};

// Block-wide exact selection without sorting: the k-th smallest (1-based, k <= n) of n uint32 keys (shared or
// global memory), found with four 8-bit histogram passes from the most significant byte (kSelThreads = 256
// threads = 256 bins).  Returns the key; n_less = how many keys are strictly smaller.  Used where only
// the VALUE of the r-th smallest matters (the keep-prefix bound qmax): a bitonic sort of the 2048-key buffer cost
// 6 600 instructions per thread, this costs a few hundred.
__device__ __forceinline__ uint32_t block_radix_select(const uint32_t* vals, int n, int k, int* hist, int* state, int tid,
                                                       int& n_less) {
    uint32_t prefix = 0;
    int kk = k;
#pragma unroll 1
    for (int shift = 24; shift >= 0; shift -= 8) {
        hist[tid] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += kSelThreads) {
            const uint32_t v = vals[i];
            if (shift == 24 || (v >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&hist[(v >> shift) & 255u], 1);
        }
        __syncthreads();
        if (tid < 32) {   // warp 0: lane l owns bins 8l .. 8l+7
            int c[8], mine = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) { c[i] = hist[tid * 8 + i]; mine += c[i]; }
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, incl, o);
                if (tid >= o) incl += up;
            }
            const unsigned reached = __ballot_sync(0xffffffffu, incl >= kk);
            const int src = __ffs(reached) - 1;   // reached != 0 because kk <= keys matching the prefix
            if (tid == src) {
                int cum = incl - mine, b = 0;
                while (cum + c[b] < kk) { cum += c[b]; ++b; }
                state[0] = tid * 8 + b;
                state[1] = kk - cum;
            }
        }
        __syncthreads();
        prefix |= static_cast<uint32_t>(state[0]) << shift;
        kk = state[1];
        __syncthreads();
    }
    n_less = k - kk;
    return prefix;
}

// Streaming "r smallest float values" over a shared buffer of 2048 value slots (ping-pong halves of kSelCap),
// compacted with block_radix_select: only the multiset of values matters (ties are interchangeable).
struct BlockMinValues {
    uint32_t* cur;          // kSelCap uint32: the buffer being filled
    uint32_t* other;        // kSelCap uint32: where a compaction writes (the two swap; no runtime-indexed pointer array,
                            // which the compiler put on the stack)
    int* count;             // entries in cur
    int* hist;              // 256
    int* state;             // 2
    unsigned int* bound;    // pass iff bits < *bound
    __device__ __forceinline__ void init(int tid) {
        if (tid == 0) { *count = 0; *bound = 0xffffffffu; }
        __syncthreads();
    }
    __device__ __forceinline__ void push(uint32_t bits) {
        if (bits < *bound) cur[atomicAdd(count, 1)] = bits;
    }
    // call after every round of <= kSelCap/2 pushes (all threads); force = last round
    __device__ __forceinline__ void maybe_compact(int r, int tid, bool force) {
        __syncthreads();
        const int c = *count;
        __syncthreads();
        if ((c > kSelCap / 2 || force) && c >= r) {
            int n_less;
            const uint32_t b = block_radix_select(cur, c, r, hist, state, tid, n_less);
            __syncthreads();
            if (tid == 0) *count = 0;
            __syncthreads();
            uint32_t* dst = other;
            for (int i = tid; i < c; i += kSelThreads) {
                const uint32_t v = cur[i];
                if (v < b) dst[atomicAdd(count, 1)] = v;
            }
            __syncthreads();
            for (int i = n_less + tid; i < r; i += kSelThreads) dst[i] = b;   // the r-th value and its ties
            __syncthreads();
            if (tid == 0) { *count = r; *bound = b; }
            other = cur;
            cur = dst;
            __syncthreads();
        }
    }
};

// ---- coarse assignment ------------------------------------------------------------------------
// d(q,c) = sum_i fma(diff_i, diff_i, .) sequentially over the dimension (the FIXED statement of
// find_k_neighbors, neighbors.cpp:30-76 with the :64 stride bug removed); the ma smallest under
// (d, c) ascending.  Two kernels: a shared-memory / register tiled distance kernel (64 queries x
// 128 centroids per CTA, every (q,c) sum still strictly in dimension order, so results are
// bit-identical to the oracle) and a per-query streaming selection.
constexpr int kCoarseTQ = 128;   // queries per CTA
constexpr int kCoarseTC = 128;   // centroids per CTA
constexpr int kCoarseTD = 16;    // dimensions per shared-memory slab

// 16 x 16 threads; thread (ty, tx) owns queries 8*ty..8*ty+7 and centroids 8*tx..8*tx+7: an 8 x 8 register tile, fed
// per dimension by four 128-bit shared-memory loads (the slabs are stored dimension-major, [d][query] and
// [d][centroid], so a thread's 8 operands are contiguous; the query loads are warp broadcasts, the centroid loads
// of the 16 tx values cover 512 contiguous bytes) for 64 sub + 64 fma.  The direct form costs two FP32 operations
// per element and cannot use the tensor cores, but it is the arithmetic the oracle pins: every (q, c) sum runs
// strictly in dimension order.  (The 4 x 8 tile with scalar loads this replaces ran at 58 % of the FP32 pipe.)
__global__ void __launch_bounds__(256, 2) coarse_dist_kernel(const float* __restrict__ queries, int nq, int dim,
                                                             const float* __restrict__ centroids, int K,
                                                             float* __restrict__ dist) {   // [nq][K]
    __shared__ __align__(16) float sq[kCoarseTD][kCoarseTQ + 4];
    __shared__ __align__(16) float sc[kCoarseTD][kCoarseTC + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int q0 = blockIdx.x * kCoarseTQ, c0 = blockIdx.y * kCoarseTC;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[i][k] = 0.f;
    for (int d0 = 0; d0 < dim; d0 += kCoarseTD) {
        const int dn = min(kCoarseTD, dim - d0);
        __syncthreads();
        // stage: consecutive threads read consecutive dimensions of one row (coalesced), store transposed
        for (int i = tid; i < kCoarseTQ * kCoarseTD; i += 256) {
            const int r = i / kCoarseTD, c = i % kCoarseTD;
            sq[c][r] = (q0 + r < nq && c < dn) ? queries[static_cast<size_t>(q0 + r) * dim + d0 + c] : 0.f;
        }
        for (int i = tid; i < kCoarseTC * kCoarseTD; i += 256) {
            const int r = i / kCoarseTD, c = i % kCoarseTD;
            sc[c][r] = (c0 + r < K && c < dn) ? __ldg(centroids + static_cast<size_t>(c0 + r) * dim + d0 + c) : 0.f;
        }
        __syncthreads();
#pragma unroll 4
        for (int d = 0; d < dn; ++d) {
            const float4 xa = *reinterpret_cast<const float4*>(&sq[d][8 * ty]), xb = *reinterpret_cast<const float4*>(&sq[d][8 * ty + 4]);
            const float4 ya = *reinterpret_cast<const float4*>(&sc[d][8 * tx]), yb = *reinterpret_cast<const float4*>(&sc[d][8 * tx + 4]);
            const float x[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
            const float y[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float diff = __fsub_rn(x[i], y[k]);
                    acc[i][k] = __fmaf_rn(diff, diff, acc[i][k]);
                }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int q = q0 + 8 * ty + i;
        if (q < nq) {
            float* row = dist + static_cast<size_t>(q) * K + c0 + 8 * tx;
            if (c0 + 8 * tx + 7 < K && (K & 3) == 0) {
                *reinterpret_cast<float4*>(row) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
                *reinterpret_cast<float4*>(row + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (c0 + 8 * tx + k < K) row[k] = acc[i][k];
            }
        }
    }
}

// out_assign (cell indices) or out_keys (distance bits << 32 | index_base + cell, ~0 padded) is written.
__global__ void __launch_bounds__(kSelThreads) coarse_select_kernel(const float* __restrict__ dist, int K, int ma,
                                                                    int32_t* __restrict__ out_assign,
                                                                    uint64_t* __restrict__ out_keys, uint32_t index_base) {
    __shared__ uint64_t keys[kSelCap];
    __shared__ int count;
    __shared__ unsigned long long bound_key;
    const int q = blockIdx.x, tid = threadIdx.x;
    const float* d = dist + static_cast<size_t>(q) * K;
    BlockTopK top{keys, &count, &bound_key};
    top.init(tid);
    auto key_of = [&](int c) { return (static_cast<uint64_t>(__float_as_uint(d[c])) << 32) | static_cast<uint32_t>(c); };
    if (ma <= kSelThreads / 2 && K >= 8 * kSelThreads) {
        // Pre-bound: the cells are dealt to the 256 threads round-robin; the ma-th smallest of the
        // 256 per-thread minima has at least ma keys at or below it, so it is a valid filter, and
        // for unordered data it lets through only a little more than ma keys.  One cheap rank
        // computation replaces the sort of the first 1024 keys; the streaming rounds below
        // (with their compaction as the safety net for adversarial orders) then see few keys.
        __shared__ __align__(16) uint64_t mins[kSelThreads];
        uint64_t mn = kEmptyKey;
        for (int c = tid; c < K; c += kSelThreads) mn = min(mn, key_of(c));
        mins[tid] = mn;
        __syncthreads();
        int rank = 0;
        for (int i = 0; i < kSelThreads; i += 2) {
            const ulonglong2 two = *reinterpret_cast<const ulonglong2*>(mins + i);   // broadcast
            rank += (two.x < mn) + (two.y < mn);
        }
        if (rank == ma - 1) bound_key = mn + 1;   // keys are distinct, so exactly one thread has this rank
        __syncthreads();
        for (int base = 0; base < K; base += kSelCap / 2) {
            for (int c = base + tid; c < min(base + kSelCap / 2, K); c += kSelThreads) top.push(key_of(c));
            top.maybe_compact(ma, tid, base + kSelCap / 2 >= K);
        }
    } else {
        // first round: nothing to filter against yet, so the keys are stored directly (no shared atomics)
        const int first = min(kSelCap / 2, K);
        for (int c = tid; c < first; c += kSelThreads) keys[c] = key_of(c);
        if (tid == 0) count = first;
        top.maybe_compact(ma, tid, first >= K);
        for (int base = first; base < K; base += kSelCap / 2) {
            for (int c = base + tid; c < min(base + kSelCap / 2, K); c += kSelThreads) top.push(key_of(c));
            top.maybe_compact(ma, tid, base + kSelCap / 2 >= K);
        }
    }
    for (int a = tid; a < ma; a += kSelThreads) {
        if (out_assign) out_assign[static_cast<size_t>(q) * ma + a] = (a < count) ? static_cast<int32_t>(static_cast<uint32_t>(keys[a])) : 0;
        if (out_keys) out_keys[static_cast<size_t>(q) * ma + a] = (a < count) ? keys[a] + index_base : ~0ull;
    }
}

// Sharded coarse assignment, second half: the ma smallest of the G*ma keys of one query
// (keys laid out [G][nq][ma], every list ascending); one CTA per query.  The largest of the lists'
// ceil(ma/G)-th keys has at least ma keys at or below it, so only keys up to it are kept (about
// ma of them instead of G*ma) and sorted in shared memory.
__global__ void __launch_bounds__(256) coarse_merge_kernel(const uint64_t* __restrict__ keys, int G, int nq, int ma,
                                                           int n_sort, int32_t* __restrict__ out_assign) {
    extern __shared__ __align__(16) uint64_t mk[];   // n_sort = pow2 >= G * ma
    __shared__ int count;
    __shared__ unsigned long long bound;
    const int q = blockIdx.x, tid = threadIdx.x;
    const int per = (ma + G - 1) / G;
    if (tid == 0) { count = 0; bound = 0ull; }
    __syncthreads();
    if (tid < G) atomicMax(&bound, static_cast<unsigned long long>(keys[(static_cast<size_t>(tid) * nq + q) * ma + per - 1]));
    for (int g = tid + 256; g < G; g += 256)
        atomicMax(&bound, static_cast<unsigned long long>(keys[(static_cast<size_t>(g) * nq + q) * ma + per - 1]));
    __syncthreads();
    const uint64_t b = bound;
    for (int i = tid; i < G * ma; i += 256) {
        const int g = i / ma, a = i - g * ma;
        const uint64_t k = keys[(static_cast<size_t>(g) * nq + q) * ma + a];
        if (k <= b && k != ~0ull) mk[atomicAdd(&count, 1)] = k;
    }
    __syncthreads();
    const int c = count;
    int n2 = 2;
    while (n2 < c) n2 <<= 1;   // <= n_sort
    for (int i = c + tid; i < n2; i += 256) mk[i] = ~0ull;
    __syncthreads();
    bitonic_sort_u64(mk, n2, tid, 256, BlockSync());
    for (int a = tid; a < ma; a += 256) {
        const uint64_t k = a < c ? mk[a] : ~0ull;
        out_assign[static_cast<size_t>(q) * ma + a] = (k == ~0ull) ? 0 : static_cast<int32_t>(static_cast<uint32_t>(k));
    }
}

// ---- residual -> rotation -> float tables: one WARP per (query, probe) ----------------------
// grid = (ceil(ma/8), nq), 8 warps per CTA.  tables[(q*ma + a)*M*16 + j*16 + c];
// tmin[q*ma + a] = min entry of that table.
// DSQ = sub-vector dimension when it is one of the common ones (loops unroll completely), 0 = any.
template <int DSQ>
__global__ void __launch_bounds__(256) tables_kernel(const float* __restrict__ queries, int dim, int M,
                                                     const float* __restrict__ codebooks,
                                                     const float* __restrict__ rotation,    // or null
                                                     const float* __restrict__ centroids,   // or null (flat)
                                                     const int32_t* __restrict__ assign, int ma,
                                                     float* __restrict__ tables, float* __restrict__ tmin, int ncl,
                                                     const uint32_t* __restrict__ part_size) {
    // ncl = log2(entries per sub-quantiser): 4 for Quick ADC, 8 for the plain 8-bit ADC tables
    // part_size != null ("owner computes", sharded inverted lists): only probes whose list lives on this device get a
    // table; the others report tmin = FLT_MAX and leave their slot untouched.
    extern __shared__ __align__(16) float sm[];   // 8 warps x 2 x dim
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.y, a_i = blockIdx.x * 8 + warp;
    if (a_i >= ma) return;
    const size_t qa = static_cast<size_t>(q) * ma + a_i;
    if (part_size && part_size[assign[qa]] == 0) {
        if (lane == 0) tmin[qa] = 3.402823466e+38f;
        return;
    }
    float* res = sm + static_cast<size_t>(warp) * 2 * dim;
    float* rot = res + dim;
    const float* query = queries + static_cast<size_t>(q) * dim;
    const float* cent = centroids ? centroids + static_cast<size_t>(assign[qa]) * dim : nullptr;
    for (int i = lane; i < dim; i += 32) res[i] = cent ? __fsub_rn(query[i], __ldg(cent + i)) : query[i];
    __syncwarp();
    const float* x = res;
    if (rotation) {
        for (int j = lane; j < dim; j += 32) {
            const float* row = rotation + static_cast<size_t>(j) * dim;
            float s = 0.f;
            for (int k = 0; k < dim; ++k) s = __fmaf_rn(res[k], __ldg(row + k), s);
            rot[j] = s;
        }
        __syncwarp();
        x = rot;
    }
    const int dsq = DSQ ? DSQ : dim / M, blocks = dsq / 8, rem = dsq % 8;
    float local_min = 3.402823466e+38f;
    const int entries = M << ncl;
    for (int e = lane; e < entries; e += 32) {
        const int j = e >> ncl;
        const float* a = x + j * dsq;
        const float* b = codebooks + static_cast<size_t>(e) * dsq;   // (j * 2^ncl + c) * dsq
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int bl = 0; bl < blocks; ++bl) {
            float av[8], bv[8];
            if ((dsq & 3) == 0) {   // 16-byte aligned rows: two 128-bit loads each (same values, same order)
                const float4 a0 = *reinterpret_cast<const float4*>(a + bl * 8), a1 = *reinterpret_cast<const float4*>(a + bl * 8 + 4);
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + bl * 8)), b1 = __ldg(reinterpret_cast<const float4*>(b + bl * 8 + 4));
                av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w; av[4] = a1.x; av[5] = a1.y; av[6] = a1.z; av[7] = a1.w;
                bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w; bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
            } else {
#pragma unroll
                for (int l = 0; l < 8; ++l) { av[l] = a[bl * 8 + l]; bv[l] = __ldg(b + bl * 8 + l); }
            }
#pragma unroll
            for (int l = 0; l < 8; ++l) {
                const float diff = __fsub_rn(av[l], bv[l]);
                acc[l] = __fmaf_rn(diff, diff, acc[l]);
            }
        }
        // reduceadd (distances.hpp:28-36)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = __fadd_rn(acc[i], acc[i + 4]);
#pragma unroll
        for (int i = 0; i < 2; ++i) acc[i] = __fadd_rn(acc[i], acc[i + 2]);
        float norm = __fadd_rn(acc[0], acc[1]);
#pragma unroll
        for (int i = 0; i < rem; ++i) {
            const float diff = __fsub_rn(__ldg(b + blocks * 8 + i), a[blocks * 8 + i]);
            norm = __fmaf_rn(diff, diff, norm);
        }
        tables[qa * entries + e] = norm;
        local_min = fminf(local_min, norm);
    }
    for (int o = 16; o > 0; o >>= 1) local_min = fminf(local_min, __shfl_xor_sync(0xffffffffu, local_min, o));
    if (lane == 0) tmin[qa] = local_min;
}

// ---- keep-prefix float ADC: grid = (splits, queries) --------------------------------------
// Every probed partition's prefix (row-major codes, `starts`) is scanned with that probe's
// float table: candidate = 0; += T[j][nibble_j] for j = 0..M-1 (query_common.hpp:72-80).
// Emits the r smallest float keys (value bits << 32 | running index) of this split.
struct PrefixArgs {
    const uint8_t* starts;          // row-major prefix codes, all partitions
    const uint64_t* start_off;      // [K] in vectors
    const uint32_t* start_size;     // [K]
    const int32_t* assign;          // [nq][ma]
    const float* tables;            // [nq][ma][M*16]
    int ma, r, M, nsplit;
    uint32_t* lists;                // [nq][nsplit][r] float bits, FLT_MAX-padded (nsplit > 1)
    float* qmax;                    // [nq], written directly when nsplit == 1
    // "owner computes" (sharded inverted lists, nsplit == 1): instead of qmax the kernel writes this device's share of
    // the query's bounds, local_out[q][0] = min entry of its tables (min over tmin[q][*]), local_out[q][1..r] = its r
    // smallest prefix distances (FLT_MAX-padded); the shards' shares are combined by bounds_combine_kernel.
    float* local_out = nullptr;
    const float* tmin = nullptr;    // [nq][ma]
    const uint32_t* owned_size = nullptr;   // [K] partition sizes on this device: a probe whose list is elsewhere is skipped
};

// local_out[q][0] of PrefixArgs: block-wide min over the query's per-probe table minima
__device__ __forceinline__ void write_local_min(const PrefixArgs& a, int q, int tid, float* red) {
    float mn = 3.402823466e+38f;
    for (int i = tid; i < a.ma; i += kSelThreads) mn = fminf(mn, a.tmin[static_cast<size_t>(q) * a.ma + i]);
    for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = mn;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < kSelThreads / 32; ++w) mn = fminf(mn, red[w]);
        a.local_out[static_cast<size_t>(q) * (a.r + 1)] = mn;
    }
}

template <int M>
__global__ void __launch_bounds__(kSelThreads) prefix_scan_kernel(const PrefixArgs a) {
    constexpr int CS = M / 2;
    __shared__ uint32_t vbuf[2][kSelCap];
    __shared__ float tab[M * 16];
    __shared__ int count, hist[256], state[2];
    __shared__ unsigned int bound;
    const int split = blockIdx.x, q = blockIdx.y, tid = threadIdx.x;
    BlockMinValues top{vbuf[0], vbuf[1], &count, hist, state, &bound};
    top.init(tid);
    for (int ar = 0; ar < a.ma; ++ar) {
        const int p = a.assign[static_cast<size_t>(q) * a.ma + ar];
        const uint32_t n = (a.owned_size && a.owned_size[p] == 0) ? 0u : a.start_size[p];
        const uint32_t v0 = static_cast<uint32_t>(static_cast<uint64_t>(n) * split / a.nsplit);
        const uint32_t v1 = static_cast<uint32_t>(static_cast<uint64_t>(n) * (split + 1) / a.nsplit);
        if (v1 == v0) continue;   // block-uniform
        __syncthreads();
        for (int i = tid; i < M * 16; i += kSelThreads)
            tab[i] = a.tables[(static_cast<size_t>(q) * a.ma + ar) * M * 16 + i];
        __syncthreads();
        const uint8_t* codes = a.starts + a.start_off[p] * CS;
        static_assert(kSelCap / 2 % kSelThreads == 0, "a round is a whole number of vectors per thread");
        constexpr int VPT = kSelCap / 2 / kSelThreads;   // vectors per thread and round; loads first: the prefix streams from L2 / HBM
        for (uint32_t base = v0; base < v1; base += kSelCap / 2) {
            const uint32_t vend = min(base + kSelCap / 2, v1);
            uint32_t w[VPT][CS / 4];
#pragma unroll
            for (int u = 0; u < VPT; ++u) {
                const uint32_t v = base + tid + u * kSelThreads;
                if (v < vend) {
                    if constexpr (CS == 8) {
                        const uint2 c = *reinterpret_cast<const uint2*>(codes + static_cast<size_t>(v) * CS);
                        w[u][0] = c.x; w[u][1] = c.y;
                    } else {
                        const uint4 c = *reinterpret_cast<const uint4*>(codes + static_cast<size_t>(v) * CS);
                        w[u][0] = c.x; w[u][1] = c.y; w[u][2] = c.z; w[u][3] = c.w;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < VPT; ++u) {
                if (base + tid + u * kSelThreads < vend) {
                    float s = 0.f;
#pragma unroll
                    for (int j = 0; j < M; ++j) s = __fadd_rn(s, tab[j * 16 + ((w[u][j >> 3] >> (4 * (j & 7))) & 15u)]);
                    top.push(__float_as_uint(s));   // sums of non-negative entries: the bit patterns order like the values
                }
            }
            top.maybe_compact(a.r, tid, false);
        }
    }
    top.maybe_compact(a.r, tid, true);
    // r smallest values of this split (FLT_MAX-padded); a single split needs no merge: the r-th value is qmax
    // (FLT_MAX when the prefix is too short)
    const int c = count;
    const uint32_t* res = top.cur;
    if (a.local_out) {
        float* dst = a.local_out + static_cast<size_t>(q) * (a.r + 1) + 1;
        for (int i = tid; i < a.r; i += kSelThreads) dst[i] = (i < c) ? __uint_as_float(res[i]) : 3.402823466e+38f;
        write_local_min(a, q, tid, reinterpret_cast<float*>(hist));
    } else if (a.nsplit == 1) {
        if (tid == 0) a.qmax[q] = (c >= a.r) ? __uint_as_float(bound) : 3.402823466e+38f;
    } else {
        uint32_t* dst = a.lists + (static_cast<size_t>(q) * a.nsplit + split) * a.r;
        for (int i = tid; i < a.r; i += kSelThreads) dst[i] = (i < c) ? res[i] : 0x7f7fffffu;
    }
}

// Short prefixes (inverted lists): one warp per probe, lanes over the few prefix vectors; the
// CTA keeps one top-r buffer.  grid = queries.
template <int M>
__global__ void __launch_bounds__(kSelThreads) prefix_scan_probes_kernel(const PrefixArgs a) {
    constexpr int CS = M / 2;
    __shared__ uint64_t keys[kSelCap];
    __shared__ float tab[8][M * 16];
    __shared__ int count;
    __shared__ unsigned long long bound_key;
    const int q = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    BlockTopK top{keys, &count, &bound_key};
    top.init(tid);
    // rounds of 8 probes; every probe contributes at most kSelCap/16 = 128 prefix vectors per round
    for (int a0 = 0; a0 < a.ma; a0 += 8) {
        const int ar = a0 + warp;
        uint32_t n = 0;
        const uint8_t* codes = nullptr;
        if (ar < a.ma) {
            const int p = a.assign[static_cast<size_t>(q) * a.ma + ar];
            n = (a.owned_size && a.owned_size[p] == 0) ? 0u : a.start_size[p];
            codes = a.starts + a.start_off[p] * CS;
            if (n)
                for (int i = lane; i < M * 16; i += 32) tab[warp][i] = a.tables[(static_cast<size_t>(q) * a.ma + ar) * M * 16 + i];
        }
        __syncwarp();
        for (uint32_t v = lane; v < n; v += 32) {   // n <= 128 (host guarantees max_start <= 128 for this kernel)
            uint32_t w[CS / 4];
            if constexpr (CS == 8) {
                const uint2 c = *reinterpret_cast<const uint2*>(codes + static_cast<size_t>(v) * CS);
                w[0] = c.x; w[1] = c.y;
            } else {
                const uint4 c = *reinterpret_cast<const uint4*>(codes + static_cast<size_t>(v) * CS);
                w[0] = c.x; w[1] = c.y; w[2] = c.z; w[3] = c.w;
            }
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < M; ++j) s = __fadd_rn(s, tab[warp][j * 16 + ((w[j >> 3] >> (4 * (j & 7))) & 15u)]);
            top.push((static_cast<uint64_t>(__float_as_uint(s)) << 32) | (static_cast<uint32_t>(ar) << 8) | v);
        }
        top.maybe_compact(a.r, tid, a0 + 8 >= a.ma);
    }
    if (a.local_out) {   // the list is sorted after the forced compaction; unused slots hold kEmptyKey
        float* dst = a.local_out + static_cast<size_t>(q) * (a.r + 1) + 1;
        const int c = count;
        for (int i = tid; i < a.r; i += kSelThreads)
            dst[i] = (i < c) ? __uint_as_float(static_cast<uint32_t>(keys[i] >> 32)) : 3.402823466e+38f;
        write_local_min(a, q, tid, &tab[0][0]);
    } else if (tid == 0) {
        a.qmax[q] = (count == a.r) ? __uint_as_float(static_cast<uint32_t>(keys[a.r - 1] >> 32)) : 3.402823466e+38f;
    }
}

// ---- "owner computes": the shards' bound shares -> the query's qmin / qmax ---------------------------------------
// gathered[g][q][0] = min table entry on shard g, [1..r] = its r smallest prefix distances (FLT_MAX-padded).
// qmin = min over shards (the min over all ma tables, db_query_4.cpp:256-260), qmax = the r-th smallest of the union
// (= the r-th smallest prefix distance over all probed lists, query_scan_start, db_query_4.cpp:230-242): both are
// order-independent selections of values every shard computed with the unsharded arithmetic, so they are bit-identical
// to the unsharded bounds.  grid = queries, dynamic shared memory G * r * 4 bytes.
__global__ void __launch_bounds__(kSelThreads) bounds_combine_kernel(const float* __restrict__ gathered, int G, int nq, int r,
                                                                    float* __restrict__ qmin_raw, float* __restrict__ qmax) {
    extern __shared__ uint32_t cvals[];
    __shared__ int hist[256], state[2];
    const int q = blockIdx.x, tid = threadIdx.x;
    float mn = 3.402823466e+38f;
    for (int g = 0; g < G; ++g) {
        const float* src = gathered + (static_cast<size_t>(g) * nq + q) * (r + 1);
        mn = fminf(mn, src[0]);
        for (int i = tid; i < r; i += kSelThreads) cvals[g * r + i] = __float_as_uint(src[1 + i]);
    }
    __syncthreads();
    int n_less;
    const uint32_t b = block_radix_select(cvals, G * r, r, hist, state, tid, n_less);
    if (tid == 0) { qmin_raw[q] = mn; qmax[q] = __uint_as_float(b); }   // FLT_MAX when fewer than r prefix vectors exist
}

// (int) trunc of the correctly rounded quotient t / delta (what QuantizerMAX computes, db_query_4.cpp:44-55) for
// 0 <= t < qmax - qmin, i.e. a quotient below 128.  t * (1/delta) is within ~2e-5 of the true quotient there, so unless
// it falls within 1e-3 of an integer its floor IS the result; only those rare entries pay for the IEEE division.
__device__ __forceinline__ int quantize_entry(float t, float delta, float inv_delta) {
    const float p = __fmul_rn(t, inv_delta);
    const float fl = floorf(p);
    const float frac = __fsub_rn(p, fl);
    if (frac > 1e-3f && frac < 0.999f) return static_cast<int>(fl);
    return static_cast<int>(__fdiv_rn(t, delta));
}

// ---- inverted lists: the whole per-query table pipeline in ONE kernel, tables resident in shared memory ----------
// grid = queries, 256 threads.  For one query: residual -> (rotation) -> float tables of all ma probes (shared memory,
// ma * M * 64 bytes) -> float ADC of the probes' keep-prefixes against them, r-th smallest = qmax (block radix
// select) -> qmin = min entry (clamped at 0) -> int8 tables (written once to global memory for the scan, kept in
// shared memory too) -> int8 distances of the same prefix vectors -> the query's shared bound (r-th smallest).
// Replaces tables_kernel + prefix_scan(_probes)_kernel + quantize_kernel + prefix_hist_kernel + prefix_bound_kernel and
// their round trips of the float tables through HBM (config 5: 3.9 GB per 10 000-query batch).  Every float operation
// is the one the separate kernels (and oracle/qadc_oracle.c) perform, in the same order: bit-identical results.
struct IvfPrepArgs {
    const float* queries; int dim;
    const float* codebooks; const float* rotation; const float* centroids;
    const int32_t* assign; int ma, r;
    const uint8_t* starts; const uint64_t* start_off; const uint32_t* start_size;
    float* tables_out;      // [nq][ma][M*16] or null
    int8_t* qtables;        // [nq][ma][M*16]
    float* qmin; float* qmax; int* shared_bound; int* err;
};

__host__ __device__ inline size_t ivf_prep_smem_bytes(int m, int ma, int dim) {
    size_t b = static_cast<size_t>(ma) * m * 16 * 4;          // float tables, later the int8 tables in their first quarter
    b += static_cast<size_t>(8) * 2 * dim * 4;                // per-warp residual + rotated copy
    b += static_cast<size_t>(ma + 1) * 4 + static_cast<size_t>(ma) * 8;   // prefix offsets, first prefix vector of each probe
    b += 2 * kSelCap * 4;                                     // value buffers of the radix select
    b += static_cast<size_t>(ma) * m * 16;                    // int8 tables (kept for the shared-bound seed)
    return b + 64;
}

template <int M, int DSQ>
__global__ void __launch_bounds__(256) ivf_prepare_kernel(const IvfPrepArgs a) {
    constexpr int CS = M / 2, TE = M * 16;
    extern __shared__ __align__(16) uint8_t ps[];
    float* tab = reinterpret_cast<float*>(ps);                                   // [ma][TE]
    float* resbuf = tab + static_cast<size_t>(a.ma) * TE;                         // [8][2*dim]
    uint64_t* pfirst = reinterpret_cast<uint64_t*>(resbuf + 16 * a.dim);          // [ma] (8-byte aligned: TE, dim*16 floats)
    int* poff = reinterpret_cast<int*>(pfirst + a.ma);                           // [ma + 1]
    uint32_t* vb0 = reinterpret_cast<uint32_t*>(poff + a.ma + 1);
    uint32_t* vb1 = vb0 + kSelCap;
    int8_t* itab = reinterpret_cast<int8_t*>((reinterpret_cast<uintptr_t>(vb1 + kSelCap) + 15) & ~static_cast<uintptr_t>(15));   // [ma][TE]
    __shared__ int count, hist[256], state[2];
    __shared__ unsigned int bound;
    __shared__ float red[8];
    const int q = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dim = a.dim, ma = a.ma;
    const float* query = a.queries + static_cast<size_t>(q) * dim;

    // ---- 1. float tables (the arithmetic of tables_kernel) ----
    // One warp per probe, no block barrier: lane l owns the table entries e = l, l + 32, ... (sub-quantiser e >> 4,
    // centroid e & 15) of every probe its warp handles, so (for the specialised sub-vector sizes) its codebook rows
    // live in registers for the whole query; the probe's residual sits in the warp's shared-memory slot; the next
    // probe's centroid components are prefetched while the current entries are computed.
    float local_min = 3.402823466e+38f;
    {
        constexpr int EPL = TE / 32;                  // entries per lane and probe: 8 (16x4) or 16 (32x4)
        constexpr bool kRegCb = DSQ != 0 && EPL * DSQ <= 64;
        constexpr int CBR = kRegCb ? DSQ : 1;
        const int dsq = DSQ ? DSQ : dim / M, blocks = dsq / 8, rem = dsq % 8;
        float cbr[kRegCb ? EPL : 1][CBR];
        if constexpr (kRegCb) {
#pragma unroll
            for (int k = 0; k < EPL; ++k)
#pragma unroll
                for (int l = 0; l < DSQ; ++l) cbr[k][l] = __ldg(a.codebooks + static_cast<size_t>(lane + 32 * k) * DSQ + l);
        }
        float* res = resbuf + static_cast<size_t>(warp) * 2 * dim;
        float* rot = res + dim;
        const int32_t* my_assign = a.assign + static_cast<size_t>(q) * ma;
        constexpr int PF = 8;                          // prefetched centroid components per lane (dim <= 256)
        const bool pf = dim <= 32 * PF;
        float qv[PF], cn[PF];
#pragma unroll
        for (int i = 0; i < PF; ++i) {
            const int d = lane + 32 * i;
            qv[i] = (pf && d < dim) ? query[d] : 0.f;
            cn[i] = (pf && d < dim && warp < ma) ? __ldg(a.centroids + static_cast<size_t>(my_assign[warp]) * dim + d) : 0.f;
        }
        for (int a_i = warp; a_i < ma; a_i += 8) {
            const size_t qa = static_cast<size_t>(q) * ma + a_i;
            __syncwarp();   // every lane is done with the previous probe's residual
            if (pf) {
#pragma unroll
                for (int i = 0; i < PF; ++i) {
                    const int d = lane + 32 * i;
                    if (d < dim) res[d] = __fsub_rn(qv[i], cn[i]);
                }
                if (a_i + 8 < ma) {
                    const float* cnext = a.centroids + static_cast<size_t>(my_assign[a_i + 8]) * dim;
#pragma unroll
                    for (int i = 0; i < PF; ++i) {
                        const int d = lane + 32 * i;
                        if (d < dim) cn[i] = __ldg(cnext + d);
                    }
                }
            } else {
                const float* cent = a.centroids + static_cast<size_t>(my_assign[a_i]) * dim;
                for (int i = lane; i < dim; i += 32) res[i] = __fsub_rn(query[i], __ldg(cent + i));
            }
            __syncwarp();
            const float* x = res;
            if (a.rotation) {
                for (int j = lane; j < dim; j += 32) {
                    const float* row = a.rotation + static_cast<size_t>(j) * dim;
                    float s = 0.f;
                    for (int k = 0; k < dim; ++k) s = __fmaf_rn(res[k], __ldg(row + k), s);
                    rot[j] = s;
                }
                __syncwarp();
                x = rot;
            }
#pragma unroll
            for (int k = 0; k < EPL; ++k) {
                const int e = lane + 32 * k;
                const float* xa = x + (e >> 4) * dsq;
                const float* b = a.codebooks + static_cast<size_t>(e) * dsq;
                float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int bl = 0; bl < blocks; ++bl) {
                    float av[8];
                    if ((dsq & 3) == 0) {   // 16-byte aligned sub-vectors: two 128-bit shared-memory loads (same values, same order)
                        const float4 a0 = *reinterpret_cast<const float4*>(xa + bl * 8), a1 = *reinterpret_cast<const float4*>(xa + bl * 8 + 4);
                        av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w; av[4] = a1.x; av[5] = a1.y; av[6] = a1.z; av[7] = a1.w;
                    } else {
#pragma unroll
                        for (int l = 0; l < 8; ++l) av[l] = xa[bl * 8 + l];
                    }
#pragma unroll
                    for (int l = 0; l < 8; ++l) {
                        const float bv = kRegCb ? cbr[kRegCb ? k : 0][(bl * 8 + l) % CBR] : __ldg(b + bl * 8 + l);
                        const float diff = __fsub_rn(av[l], bv);
                        acc[l] = __fmaf_rn(diff, diff, acc[l]);
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i] = __fadd_rn(acc[i], acc[i + 4]);
#pragma unroll
                for (int i = 0; i < 2; ++i) acc[i] = __fadd_rn(acc[i], acc[i + 2]);
                float norm = __fadd_rn(acc[0], acc[1]);
#pragma unroll
                for (int i = 0; i < rem; ++i) {
                    const float bv = kRegCb ? cbr[kRegCb ? k : 0][(blocks * 8 + i) % CBR] : __ldg(b + blocks * 8 + i);
                    const float diff = __fsub_rn(bv, xa[blocks * 8 + i]);
                    norm = __fmaf_rn(diff, diff, norm);
                }
                tab[static_cast<size_t>(a_i) * TE + e] = norm;
                if (a.tables_out) a.tables_out[qa * TE + e] = norm;
                local_min = fminf(local_min, norm);
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) local_min = fminf(local_min, __shfl_xor_sync(0xffffffffu, local_min, o));
    if (lane == 0) red[warp] = local_min;
    // ---- 2. where each probe's keep-prefix starts in the flattened item space ----
    if (warp == 0) {
        int running = 0;
        for (int base = 0; base < ma; base += 32) {
            const int i = base + lane;
            int n = 0;
            if (i < ma) {
                const int p = a.assign[static_cast<size_t>(q) * ma + i];
                n = static_cast<int>(a.start_size[p]);
                pfirst[i] = a.start_off[p];
            }
            int incl = n;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (i < ma) poff[i + 1] = running + incl;
            running += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) poff[0] = 0;
    }
    BlockMinValues top{vb0, vb1, &count, hist, state, &bound};
    top.init(tid);   // (a block barrier: tables, red[], poff[] are visible from here on)
    const int total = poff[ma];
    auto locate = [&](int item, int& probe, uint32_t (&w)[CS / 4]) {   // probe of a flattened item and its code words
        int lo = 0, hi = ma;   // poff[lo] <= item < poff[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (poff[mid] <= item) lo = mid; else hi = mid;
        }
        probe = lo;
        const uint8_t* c = a.starts + (pfirst[lo] + static_cast<uint32_t>(item - poff[lo])) * CS;
        if constexpr (CS == 8) {
            const uint2 v = *reinterpret_cast<const uint2*>(c);
            w[0] = v.x; w[1] = v.y;
        } else {
            const uint4 v = *reinterpret_cast<const uint4*>(c);
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        }
    };
    // ---- 3. float ADC of the prefixes (scan_4, query_common.hpp:59-90) -> qmax = r-th smallest ----
    for (int base = 0; base < total; base += kSelCap / 2) {
        for (int item = base + tid; item < min(base + kSelCap / 2, total); item += kSelThreads) {
            int probe;
            uint32_t w[CS / 4];
            locate(item, probe, w);
            const float* t = tab + static_cast<size_t>(probe) * TE;
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < M; ++j) s = __fadd_rn(s, t[j * 16 + ((w[j >> 3] >> (4 * (j & 7))) & 15u)]);
            top.push(__float_as_uint(s));
        }
        top.maybe_compact(a.r, tid, false);
    }
    top.maybe_compact(a.r, tid, true);
    const float qmax = (count >= a.r) ? __uint_as_float(bound) : 3.402823466e+38f;
    // ---- 4. bounds + int8 tables (QuantizerMAX, db_query_4.cpp:37-71, :256-284) ----
    float qmin = red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) qmin = fminf(qmin, red[w]);
    if (qmin < 0.f) qmin = 0.f;
    if (tid == 0) {
        a.qmin[q] = qmin;
        a.qmax[q] = qmax;
        if (qmax > 1e30f) atomicExch(a.err, 1);
    }
    const float delta = __fdiv_rn(__fsub_rn(qmax, qmin), 127.0f);
    const float inv_delta = __fdiv_rn(1.0f, delta);
    const int entries = ma * TE;
    int8_t* gq = a.qtables + static_cast<size_t>(q) * entries;
    for (int e = 4 * tid; e < entries; e += 4 * kSelThreads) {   // own shared-memory region: no barrier per chunk
        const float4 v = *reinterpret_cast<const float4*>(tab + e);
        const float vv[4] = {v.x, v.y, v.z, v.w};
        int8_t qv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float x = vv[i] < 0.f ? 0.f : vv[i];
            qv[i] = (x >= qmax) ? static_cast<int8_t>(127) : static_cast<int8_t>(quantize_entry(__fsub_rn(x, qmin), delta, inv_delta));
        }
        const char4 out = make_char4(qv[0], qv[1], qv[2], qv[3]);
        *reinterpret_cast<char4*>(itab + e) = out;
        *reinterpret_cast<char4*>(gq + e) = out;
    }
    // ---- 5. the query's shared bound: r-th smallest int8 distance among the same prefix vectors ----
    hist[tid] = 0;
    __syncthreads();
    for (int item = tid; item < total; item += kSelThreads) {
        int probe;
        uint32_t w[CS / 4];
        locate(item, probe, w);
        const int8_t* t = itab + static_cast<size_t>(probe) * TE;
        int sum = 0;
#pragma unroll
        for (int j = 0; j < M; ++j) sum += t[j * 16 + ((w[j >> 3] >> (4 * (j & 7))) & 15u)];
        atomicAdd(&hist[min(sum, 127)], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int cum = 0, b = 126;   // fewer than r prefix vectors below 127: everything below 127 may pass
        for (int d = 0; d < 127; ++d) {
            cum += hist[d];
            if (cum >= a.r) { b = d; break; }
        }
        a.shared_bound[q] = b;
    }
}

// ---- PQ encoder ("next" row N1) ---------------------------------------------------------------
// base_pq::encode_multiple_vectors + multiple_set_bits_4 / _native<uint8_t> (quantizers.hpp:36-68,
// :222-245): one thread per (vector, code byte) finds the nearest of the 2^bits centroids (direct
// squared distance, first minimum wins like the k=1 heap) for the sub-quantisers of that byte:
// 4-bit codes hold idx[2b] | idx[2b+1]<<4, 8-bit codes idx[b], 16-bit codes idx[j] as two bytes.
// `centroids`/`assign` given: encode the residual x - centroid[assign[v]] (index_db::add_vectors).
__global__ void __launch_bounds__(256) encode_kernel(const float* __restrict__ vectors, uint32_t count, int dim, int M,
                                                     int bits, const float* __restrict__ codebooks,
                                                     const float* __restrict__ centroids,
                                                     const int32_t* __restrict__ assign, uint8_t* __restrict__ codes) {
    const int CS = M * bits / 8, dsq = dim / M, per = 8 / bits, ncent = 1 << bits;
    const size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (bits == 16) {
        // one thread per (vector, sub-quantiser): the nearest of 65 536 centroids, stored as a little-endian uint16
        // (multiple_set_bits_native<std::uint16_t>)
        if (t >= static_cast<size_t>(count) * M) return;
        const uint32_t v = static_cast<uint32_t>(t / M);
        const int j = static_cast<int>(t % M);
        const float* x = vectors + static_cast<size_t>(v) * dim;
        const float* cent = centroids ? centroids + static_cast<size_t>(assign[v]) * dim : nullptr;
        int best = 0;
        float bd = 0.f;
        for (int c = 0; c < ncent; ++c) {
            const float* cb = codebooks + (static_cast<size_t>(j) * ncent + c) * dsq;
            float s = 0.f;
            for (int i = 0; i < dsq; ++i) {
                const float xi = cent ? __fsub_rn(x[j * dsq + i], __ldg(cent + j * dsq + i)) : x[j * dsq + i];
                const float diff = __fsub_rn(xi, __ldg(cb + i));
                s = __fmaf_rn(diff, diff, s);
            }
            if (c == 0 || s < bd) { bd = s; best = c; }
        }
        codes[static_cast<size_t>(v) * CS + 2 * j] = static_cast<uint8_t>(best & 255);
        codes[static_cast<size_t>(v) * CS + 2 * j + 1] = static_cast<uint8_t>(best >> 8);
        return;
    }
    if (t >= static_cast<size_t>(count) * CS) return;
    const uint32_t v = static_cast<uint32_t>(t / CS);
    const int b = static_cast<int>(t % CS);
    const float* x = vectors + static_cast<size_t>(v) * dim;
    const float* cent = centroids ? centroids + static_cast<size_t>(assign[v]) * dim : nullptr;
    uint32_t code = 0;
    for (int h = 0; h < per; ++h) {
        const int j = per * b + h;
        int best = 0;
        float bd = 0.f;
        for (int c = 0; c < ncent; ++c) {
            const float* cb = codebooks + (static_cast<size_t>(j) * ncent + c) * dsq;
            float s = 0.f;
            for (int i = 0; i < dsq; ++i) {
                const float xi = cent ? __fsub_rn(x[j * dsq + i], __ldg(cent + j * dsq + i)) : x[j * dsq + i];
                const float diff = __fsub_rn(xi, __ldg(cb + i));
                s = __fmaf_rn(diff, diff, s);
            }
            if (c == 0 || s < bd) { bd = s; best = c; }
        }
        code |= static_cast<uint32_t>(best) << (bits * h);
    }
    codes[t] = static_cast<uint8_t>(code);
}

// opq::rotate_multiple_vectors (quantizers.hpp:289-301): out = X * R^T, sequential fma over k.  With `centroids` /
// `assign` the residual x - centroid[assign[v]] is what is rotated (index_db::add_vectors with an opq:
// assign_single_compute_residuals, then encode_multiple_vectors rotates, databases.hpp:252-298, quantizers.hpp:222-224).
__global__ void __launch_bounds__(256) rotate_kernel(const float* __restrict__ vectors, uint32_t count, int dim,
                                                     const float* __restrict__ rotation, const float* __restrict__ centroids,
                                                     const int32_t* __restrict__ assign, float* __restrict__ out) {
    const size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= static_cast<size_t>(count) * dim) return;
    const size_t v = t / dim;
    const int j = static_cast<int>(t % dim);
    const float* x = vectors + v * dim;
    const float* row = rotation + static_cast<size_t>(j) * dim;
    const float* cent = centroids ? centroids + static_cast<size_t>(assign[v]) * dim : nullptr;
    float s = 0.f;
    if (cent) for (int k = 0; k < dim; ++k) s = __fmaf_rn(__fsub_rn(x[k], __ldg(cent + k)), __ldg(row + k), s);
    else for (int k = 0; k < dim; ++k) s = __fmaf_rn(x[k], __ldg(row + k), s);
    out[t] = s;
}

// ---- bounds + quantisation: one CTA per query ------------------------------------------------
// qmin = min over all ma tables (negatives clamped to 0, also in `tables`), qmax from the
// prefix scan; q(v) = 127 if v >= qmax else (int8) trunc((v - qmin) / delta),
// delta = (qmax - qmin) / 127.  err[0] is set when qmax > 1e30 (db_query_4.cpp:271-274).
// One query's share (all 256 threads of its CTA call it): `red` = 8 floats of shared memory.
__device__ __forceinline__ void quantize_query(float* __restrict__ tables, const float* __restrict__ tmin, const float qmax,
                                               int ma, int M, int8_t* __restrict__ qtables, float* __restrict__ qmin_out,
                                               int* __restrict__ err, const float* __restrict__ qmin_in,
                                               const int32_t* __restrict__ assign, const uint32_t* __restrict__ part_size,
                                               const int q, const int tid, float* red) {
    float mn = qmin_in ? qmin_in[q] : 3.402823466e+38f;
    if (!qmin_in)
        for (int a = tid; a < ma; a += 256) mn = fminf(mn, tmin[static_cast<size_t>(q) * ma + a]);
    for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    if ((tid & 31) == 0) red[tid >> 5] = mn;
    __syncthreads();
    float qmin = red[0];
    for (int w = 1; w < 8; ++w) qmin = fminf(qmin, red[w]);
    const bool clamp = qmin < 0.f;
    if (clamp) qmin = 0.f;
    if (tid == 0) {
        qmin_out[q] = qmin;
        if (qmax > 1e30f) atomicExch(err, 1);
    }
    const float delta = __fdiv_rn(__fsub_rn(qmax, qmin), 127.0f);
    const size_t base = static_cast<size_t>(q) * ma * M * 16;
    const int total = ma * M * 16;
    for (int e = tid * 4; e < total; e += 256 * 4) {   // M*16 is a multiple of 4
        if (part_size && part_size[assign[static_cast<size_t>(q) * ma + e / (M * 16)]] == 0) continue;
        float4 v = *reinterpret_cast<float4*>(tables + base + e);
        float vv[4] = {v.x, v.y, v.z, v.w};
        char4 out;
        int8_t qv[4];
        bool wrote = false;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (clamp && vv[i] < 0.f) { vv[i] = 0.f; wrote = true; }
            qv[i] = (vv[i] >= qmax) ? static_cast<int8_t>(127)
                                    : static_cast<int8_t>(static_cast<int>(__fdiv_rn(__fsub_rn(vv[i], qmin), delta)));
        }
        if (wrote) *reinterpret_cast<float4*>(tables + base + e) = make_float4(vv[0], vv[1], vv[2], vv[3]);
        out.x = qv[0]; out.y = qv[1]; out.z = qv[2]; out.w = qv[3];
        *reinterpret_cast<char4*>(qtables + base + e) = out;
    }
}

__global__ void __launch_bounds__(256) quantize_kernel(float* __restrict__ tables, const float* __restrict__ tmin,
                                                       const float* __restrict__ qmax_in, int ma, int M,
                                                       int8_t* __restrict__ qtables, float* __restrict__ qmin_out,
                                                       int* __restrict__ err, const float* __restrict__ qmin_in = nullptr,
                                                       const int32_t* __restrict__ assign = nullptr,
                                                       const uint32_t* __restrict__ part_size = nullptr,
                                                       const uint32_t* __restrict__ sel_lists = nullptr, int sel_n = 0,
                                                       int sel_r = 0, float* qmax_out = nullptr) {
    // qmin_in / assign / part_size ("owner computes"): the query's minimum comes from bounds_combine_kernel and only the
    // tables of probes whose list lives on this device exist.
    // sel_lists (flat databases, split prefix scan): qmax is selected here — the sel_r-th smallest of the sel_n values
    // the splits left (a launch of its own before) — and stored to qmax_out.
    __shared__ float red[8];
    __shared__ int sel_hist[256], sel_state[2];
    const int q = blockIdx.x, tid = threadIdx.x;
    float qmax;
    if (sel_lists) {
        int n_less;
        qmax = __uint_as_float(block_radix_select(sel_lists + static_cast<size_t>(q) * sel_n, sel_n, sel_r, sel_hist,
                                                  sel_state, tid, n_less));
        if (tid == 0) qmax_out[q] = qmax;   // FLT_MAX (the padding) when fewer than sel_r prefix vectors exist
    } else {
        qmax = qmax_in[q];
    }
    quantize_query(tables, tmin, qmax, ma, M, qtables, qmin_out, err, qmin_in, assign, part_size, q, tid, red);
}

}  // namespace qadc
