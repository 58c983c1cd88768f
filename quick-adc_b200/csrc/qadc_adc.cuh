// qadc_adc.cuh — the plain ADC scan of the reference's db_query tool ("next" row N4):
// scanner_simple + scan_standard<uint8_t,NSQ> / scan_4<NSQ> (db_query.cpp:17-46,
// query_common.hpp:59-118) over row-major codes with float tables of 2^bits entries (4-, 8- and 16-bit
// sub-quantisers: every (nsq, bits) pair get_scan_func accepts, query_common.hpp:122-147).
// Selection rule: the r smallest under (distance, probe rank, position) — what the reference's
// strict heap test yields when partitions are visited in probe order.  Distances are summed in
// sub-quantiser order with __fadd_rn (bit-identical to the oracle; the reference's -ffast-math
// build may reassociate, hence a 1e-5 tolerance against it).
#pragma once
#include "qadc_tables.cuh"

namespace qadc {

struct AdcScanArgs {
    const uint8_t* rows;        // row-major codes, partitions concatenated
    const uint64_t* row_off;    // [P + 1], in vectors
    const int32_t* assign;      // [nq][ma]
    const float* tables;        // [nq][ma][NSQ << BITS]
    int ma, r, nsplit;
    uint64_t* lists;            // [nq][nsplit][r] keys: distance bits << 32 | sequence number in probe order
};

// grid = (splits, queries): split s scans the s-th slice of every probed partition.
template <int BITS, int NSQ>
__global__ void __launch_bounds__(kSelThreads) adc_scan_kernel(const AdcScanArgs a) {
    constexpr int CS = NSQ * BITS / 8, NC = 1 << BITS, W = (CS + 3) / 4;
    constexpr bool kSmemTable = BITS <= 8;   // 16-bit quantisers: 65 536 floats per sub-quantiser stay in global memory (L2)
    __shared__ uint64_t keys[kSelCap];
    __shared__ float tab[kSmemTable ? NSQ * NC : 1];
    __shared__ int count;
    __shared__ unsigned long long bound_key;
    const int split = blockIdx.x, q = blockIdx.y, tid = threadIdx.x;
    BlockTopK top{keys, &count, &bound_key};
    top.init(tid);
    uint32_t seq0 = 0;   // sequence number of the probe's first vector
    for (int ar = 0; ar < a.ma; ++ar) {
        const int p = a.assign[static_cast<size_t>(q) * a.ma + ar];
        const uint32_t n = static_cast<uint32_t>(a.row_off[p + 1] - a.row_off[p]);
        const uint32_t v0 = static_cast<uint32_t>(static_cast<uint64_t>(n) * split / a.nsplit);
        const uint32_t v1 = static_cast<uint32_t>(static_cast<uint64_t>(n) * (split + 1) / a.nsplit);
        if (v1 > v0) {   // block-uniform
            const float* gtab = a.tables + (static_cast<size_t>(q) * a.ma + ar) * (NSQ * NC);
            if constexpr (kSmemTable) {
                __syncthreads();
                for (int i = tid; i < NSQ * NC; i += kSelThreads) tab[i] = gtab[i];
                __syncthreads();
            }
            const uint8_t* codes = a.rows + a.row_off[p] * CS;
            for (uint32_t base = v0; base < v1; base += kSelCap / 2) {
                for (uint32_t v = base + tid; v < min(base + kSelCap / 2, v1); v += kSelThreads) {
                    uint32_t w[W];
                    const uint8_t* c = codes + static_cast<size_t>(v) * CS;
                    if constexpr (CS == 4) {
                        w[0] = *reinterpret_cast<const uint32_t*>(c);
                    } else if constexpr (CS == 8) {
                        const uint2 t = *reinterpret_cast<const uint2*>(c);
                        w[0] = t.x; w[1] = t.y;
                    } else {
                        const uint4 t = *reinterpret_cast<const uint4*>(c);
                        w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
                    }
                    float s = 0.f;
#pragma unroll
                    for (int j = 0; j < NSQ; ++j) {
                        const uint32_t idx = BITS == 4 ? (w[j >> 3] >> (4 * (j & 7))) & 15u
                                           : BITS == 8 ? (w[j >> 2] >> (8 * (j & 3))) & 255u
                                                       : (w[j >> 1] >> (16 * (j & 1))) & 65535u;
                        s = __fadd_rn(s, kSmemTable ? tab[j * NC + idx] : __ldg(gtab + j * NC + idx));
                    }
                    top.push((static_cast<uint64_t>(__float_as_uint(s)) << 32) | (seq0 + v));
                }
                top.maybe_compact(a.r, tid, false);
            }
        }
        seq0 += n;
    }
    top.maybe_compact(a.r, tid, true);
    uint64_t* dst = a.lists + (static_cast<size_t>(q) * a.nsplit + split) * a.r;
    for (int i = tid; i < a.r; i += kSelThreads) dst[i] = keys[i];
}

// merged keys [nq][r] -> ids (labels of inverted lists, positions of a flat database) and float
// distances; empty slots: id 0, FLT_MAX (the reference's pre-filled heap entries, db_query.cpp:27-30).
__global__ void __launch_bounds__(256) adc_finalize_kernel(const uint64_t* __restrict__ keys, int r, int ma,
                                                           const int32_t* __restrict__ assign,
                                                           const uint64_t* __restrict__ row_off,
                                                           const uint32_t* __restrict__ labels,
                                                           uint32_t* __restrict__ out_ids, float* __restrict__ out_dists) {
    const int q = blockIdx.x;
    for (int i = threadIdx.x; i < r; i += 256) {
        const uint64_t k = keys[static_cast<size_t>(q) * r + i];
        uint32_t id = 0;
        float d = 3.402823466e+38f;
        if (k != kEmptyKey) {
            d = __uint_as_float(static_cast<uint32_t>(k >> 32));
            uint32_t seq = static_cast<uint32_t>(k);
            for (int ar = 0; ar < ma; ++ar) {
                const int p = assign[static_cast<size_t>(q) * ma + ar];
                const uint32_t n = static_cast<uint32_t>(row_off[p + 1] - row_off[p]);
                if (seq < n) {
                    id = labels ? labels[row_off[p] + seq] : seq;
                    break;
                }
                seq -= n;
            }
        }
        out_ids[static_cast<size_t>(q) * r + i] = id;
        out_dists[static_cast<size_t>(q) * r + i] = d;
    }
}

}  // namespace qadc
