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
            bitonic_sort_u64(keys, kSelCap, tid, kSelThreads, BlockSync());
            const int n = min(c, k);
            for (int i = k + tid; i < kSelCap; i += kSelThreads) keys[i] = kEmptyKey;
            if (tid == 0) { *count = n; *bound_key = (n == k) ? keys[k - 1] : kEmptyKey; }
            __syncthreads();
        }
    }
};

// ---- coarse assignment ------------------------------------------------------------------------
// d(q,c) = sum_i fma(diff_i, diff_i, .) sequentially over the dimension (the FIXED statement of
// find_k_neighbors, neighbors.cpp:30-76 with the :64 stride bug removed); the ma smallest under
// (d, c) ascending.  Two kernels: a shared-memory tiled distance kernel (32 queries x 64
// centroids per CTA step, every (q,c) sum still strictly in dimension order, so results are
// bit-identical to the oracle) and a per-query streaming selection.
constexpr int kCoarseTQ = 32;   // queries per CTA
constexpr int kCoarseTC = 64;   // centroids per tile
constexpr int kCoarseTD = 32;   // dimensions per smem slab

__global__ void __launch_bounds__(256) coarse_dist_kernel(const float* __restrict__ queries, int nq, int dim,
                                                          const float* __restrict__ centroids, int K,
                                                          float* __restrict__ dist) {   // [nq][K]
    __shared__ float sq[kCoarseTQ][kCoarseTD + 1];
    __shared__ float sc[kCoarseTC][kCoarseTD + 1];
    const int tid = threadIdx.x;
    const int tq = tid & 31;          // query inside the tile (lane)
    const int tc0 = (tid >> 5) * 8;   // 8 centroids per thread, warp-uniform -> broadcast reads of sc
    const int q0 = blockIdx.x * kCoarseTQ, c0 = blockIdx.y * kCoarseTC;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int d0 = 0; d0 < dim; d0 += kCoarseTD) {
        const int dn = min(kCoarseTD, dim - d0);
        __syncthreads();
        for (int i = tid; i < kCoarseTQ * kCoarseTD; i += 256) {
            const int r = i / kCoarseTD, c = i % kCoarseTD;
            sq[r][c] = (q0 + r < nq && c < dn) ? queries[static_cast<size_t>(q0 + r) * dim + d0 + c] : 0.f;
        }
        for (int i = tid; i < kCoarseTC * kCoarseTD; i += 256) {
            const int r = i / kCoarseTD, c = i % kCoarseTD;
            sc[r][c] = (c0 + r < K && c < dn) ? __ldg(centroids + static_cast<size_t>(c0 + r) * dim + d0 + c) : 0.f;
        }
        __syncthreads();
        for (int i = 0; i < dn; ++i) {
            const float x = sq[tq][i];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float diff = __fsub_rn(x, sc[tc0 + k][i]);
                acc[k] = __fmaf_rn(diff, diff, acc[k]);
            }
        }
    }
    if (q0 + tq < nq) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (c0 + tc0 + k < K) dist[static_cast<size_t>(q0 + tq) * K + c0 + tc0 + k] = acc[k];
    }
}

__global__ void __launch_bounds__(kSelThreads) coarse_select_kernel(const float* __restrict__ dist, int K, int ma,
                                                                    int32_t* __restrict__ out_assign) {
    __shared__ uint64_t keys[kSelCap];
    __shared__ int count;
    __shared__ unsigned long long bound_key;
    const int q = blockIdx.x, tid = threadIdx.x;
    const float* d = dist + static_cast<size_t>(q) * K;
    BlockTopK top{keys, &count, &bound_key};
    top.init(tid);
    for (int base = 0; base < K; base += kSelCap / 2) {
        for (int c = base + tid; c < min(base + kSelCap / 2, K); c += kSelThreads)
            top.push((static_cast<uint64_t>(__float_as_uint(d[c])) << 32) | static_cast<uint32_t>(c));
        top.maybe_compact(ma, tid, base + kSelCap / 2 >= K);
    }
    for (int a = tid; a < ma; a += kSelThreads)
        out_assign[static_cast<size_t>(q) * ma + a] = (a < count) ? static_cast<int32_t>(static_cast<uint32_t>(keys[a])) : 0;
}

// ---- residual -> rotation -> float tables: one CTA per (query, probe) --------------------
// tables[(q*ma + a)*M*16 + j*16 + c];  tmin[q*ma + a] = min entry of that table.
__global__ void __launch_bounds__(256) tables_kernel(const float* __restrict__ queries, int dim, int M,
                                                     const float* __restrict__ codebooks,
                                                     const float* __restrict__ rotation,    // or null
                                                     const float* __restrict__ centroids,   // or null (flat)
                                                     const int32_t* __restrict__ assign, int ma,
                                                     float* __restrict__ tables, float* __restrict__ tmin) {
    extern __shared__ __align__(16) float sm[];
    float* res = sm;          // dim
    float* rot = sm + dim;    // dim
    __shared__ float red[8];
    const int qa = blockIdx.x, q = qa / ma, tid = threadIdx.x;
    const float* query = queries + static_cast<size_t>(q) * dim;
    const float* cent = centroids ? centroids + static_cast<size_t>(assign[qa]) * dim : nullptr;
    for (int i = tid; i < dim; i += blockDim.x) res[i] = cent ? __fsub_rn(query[i], cent[i]) : query[i];
    __syncthreads();
    const float* x = res;
    if (rotation) {
        for (int j = tid; j < dim; j += blockDim.x) {
            const float* row = rotation + static_cast<size_t>(j) * dim;
            float s = 0.f;
            for (int k = 0; k < dim; ++k) s = __fmaf_rn(res[k], __ldg(row + k), s);
            rot[j] = s;
        }
        __syncthreads();
        x = rot;
    }
    const int dsq = dim / M, blocks = dsq / 8, rem = dsq % 8;
    float local_min = 3.402823466e+38f;
    for (int e = tid; e < M * 16; e += blockDim.x) {
        const int j = e >> 4;
        const float* a = x + j * dsq;
        const float* b = codebooks + static_cast<size_t>(e) * dsq;   // (j*16 + c) * dsq
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int bl = 0; bl < blocks; ++bl) {
#pragma unroll
            for (int l = 0; l < 8; ++l) {
                const float diff = __fsub_rn(a[bl * 8 + l], __ldg(b + bl * 8 + l));
                acc[l] = __fmaf_rn(diff, diff, acc[l]);
            }
        }
        // reduceadd (distances.hpp:28-36)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = __fadd_rn(acc[i], acc[i + 4]);
#pragma unroll
        for (int i = 0; i < 2; ++i) acc[i] = __fadd_rn(acc[i], acc[i + 2]);
        float norm = __fadd_rn(acc[0], acc[1]);
        for (int i = 0; i < rem; ++i) {
            const float diff = __fsub_rn(__ldg(b + blocks * 8 + i), a[blocks * 8 + i]);
            norm = __fmaf_rn(diff, diff, norm);
        }
        tables[static_cast<size_t>(qa) * M * 16 + e] = norm;
        local_min = fminf(local_min, norm);
    }
    for (int o = 16; o > 0; o >>= 1) local_min = fminf(local_min, __shfl_xor_sync(0xffffffffu, local_min, o));
    if ((tid & 31) == 0) red[tid >> 5] = local_min;
    __syncthreads();
    if (tid == 0) {
        float mn = red[0];
        for (int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w) mn = fminf(mn, red[w]);
        tmin[qa] = mn;
    }
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
    uint64_t* lists;                // [nq][nsplit][r]
};

template <int M>
__global__ void __launch_bounds__(kSelThreads) prefix_scan_kernel(const PrefixArgs a) {
    constexpr int CS = M / 2;
    __shared__ uint64_t keys[kSelCap];
    __shared__ float tab[M * 16];
    __shared__ int count;
    __shared__ unsigned long long bound_key;
    const int split = blockIdx.x, q = blockIdx.y, tid = threadIdx.x;
    BlockTopK top{keys, &count, &bound_key};
    top.init(tid);
    uint32_t running = 0;
    for (int ar = 0; ar < a.ma; ++ar) {
        const int p = a.assign[static_cast<size_t>(q) * a.ma + ar];
        const uint32_t n = a.start_size[p];
        const uint32_t v0 = static_cast<uint32_t>(static_cast<uint64_t>(n) * split / a.nsplit);
        const uint32_t v1 = static_cast<uint32_t>(static_cast<uint64_t>(n) * (split + 1) / a.nsplit);
        if (v1 == v0) continue;   // block-uniform
        __syncthreads();
        for (int i = tid; i < M * 16; i += kSelThreads)
            tab[i] = a.tables[(static_cast<size_t>(q) * a.ma + ar) * M * 16 + i];
        __syncthreads();
        const uint8_t* codes = a.starts + a.start_off[p] * CS;
        for (uint32_t base = v0; base < v1; base += kSelCap / 2) {
            for (uint32_t v = base + tid; v < min(base + kSelCap / 2, v1); v += kSelThreads) {
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
                for (int j = 0; j < M; ++j) s = __fadd_rn(s, tab[j * 16 + ((w[j >> 3] >> (4 * (j & 7))) & 15u)]);
                top.push((static_cast<uint64_t>(__float_as_uint(s)) << 32) | (running + (v - v0)));
            }
            top.maybe_compact(a.r, tid, false);
        }
        running += v1 - v0;
    }
    top.maybe_compact(a.r, tid, true);
    uint64_t* dst = a.lists + (static_cast<size_t>(q) * a.nsplit + split) * a.r;
    for (int i = tid; i < a.r; i += kSelThreads) dst[i] = keys[i];
}

// ---- bounds + quantisation: one CTA per (query, probe) ------------------------------------
// qmin = min over all ma tables (negatives clamped to 0, also in `tables`), qmax from the
// prefix scan; q(v) = 127 if v >= qmax else (int8) trunc((v - qmin) / delta),
// delta = (qmax - qmin) / 127.  err[0] is set when qmax > 1e30 (db_query_4.cpp:271-274).
__global__ void __launch_bounds__(256) quantize_kernel(float* __restrict__ tables, const float* __restrict__ tmin,
                                                       const float* __restrict__ qmax_in, int ma, int M,
                                                       int8_t* __restrict__ qtables, float* __restrict__ qmin_out,
                                                       int* __restrict__ err) {
    const int qa = blockIdx.x, q = qa / ma, tid = threadIdx.x;
    float qmin = tmin[static_cast<size_t>(q) * ma];
    for (int a = 1; a < ma; ++a) qmin = fminf(qmin, tmin[static_cast<size_t>(q) * ma + a]);
    const bool clamp = qmin < 0.f;
    if (clamp) qmin = 0.f;
    const float qmax = qmax_in[q];
    if (tid == 0 && qa == q * ma) {
        qmin_out[q] = qmin;
        if (qmax > 1e30f) atomicExch(err, 1);
    }
    const float delta = __fdiv_rn(__fsub_rn(qmax, qmin), 127.0f);
    for (int e = tid; e < M * 16; e += blockDim.x) {
        const size_t i = static_cast<size_t>(qa) * M * 16 + e;
        float v = tables[i];
        if (clamp && v < 0.f) { v = 0.f; tables[i] = 0.f; }
        int8_t qv;
        if (v >= qmax) qv = 127;
        else qv = static_cast<int8_t>(static_cast<int>(__fdiv_rn(__fsub_rn(v, qmin), delta)));
        qtables[i] = qv;
    }
}

}  // namespace qadc
