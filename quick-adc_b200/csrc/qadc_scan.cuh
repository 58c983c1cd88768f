// qadc_scan.cuh — layout, scan, distance-dump and top-r merge kernels (sm_100a).
// Replaces, on the device: interleave_partition_4 (simd_layout.hpp:55-65), scan_avx_4<16|32>
// + compare_extract_matches_sse + bh_push (simd_scan.hpp:63-187) and the result heap
// kv_binheap<unsigned,int8_t> (binheap.hpp:75-116) under the canonical selection rule.
#pragma once
#include "qadc_device.cuh"

namespace qadc {

// ------------------------------------------------------------------------------------------
// Layout: row-major codes -> nibble-plane superblocks.  One thread per (superblock, lane):
// it reads its 8 consecutive codes (8*M/2 contiguous bytes) and emits M words.  Vectors
// past the end of the partition replicate the last one, like the reference's pad lanes
// (simd_layout.hpp:46-50); the scan never selects them (position check).
// ------------------------------------------------------------------------------------------
template <int M>
__global__ void __launch_bounds__(256) transpose_codes_kernel(const uint8_t* __restrict__ codes,  // chunk, row-major
                                                              uint32_t count,        // vectors in this chunk
                                                              uint8_t* __restrict__ out,  // partition base (native)
                                                              uint32_t first) {      // multiple of 256
    constexpr int CS = M / 2;
    const uint32_t gidx = blockIdx.x * blockDim.x + threadIdx.x;   // group index inside the chunk
    const uint32_t n_groups = ((count + kSbVec - 1) / kSbVec) * 32;
    if (gidx >= n_groups) return;
    uint32_t word[M];
#pragma unroll
    for (int j = 0; j < M; ++j) word[j] = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        uint32_t v = gidx * 8 + k;
        if (v >= count) v = count - 1;
        const uint8_t* c = codes + static_cast<size_t>(v) * CS;
#pragma unroll
        for (int b = 0; b < CS; ++b) {
            const uint32_t byte = c[b];
            word[2 * b] |= (byte & 15u) << (4 * k);
            word[2 * b + 1] |= (byte >> 4) << (4 * k);
        }
    }
    const uint32_t sb = first / kSbVec + gidx / 32, lane = gidx % 32;
    uint4* dst = reinterpret_cast<uint4*>(out + static_cast<size_t>(sb) * sb_bytes(M));
#pragma unroll
    for (int q = 0; q < M / 4; ++q)
        dst[q * 32 + lane] = make_uint4(word[4 * q], word[4 * q + 1], word[4 * q + 2], word[4 * q + 3]);
}

// Same for a run of partitions [p0, p1) whose row-major codes lie back to back in `codes` (partition p0
// first): one launch for thousands of short inverted lists instead of one per list.  One thread per
// (superblock, lane) of the run; the partition of a superblock is found by bisection of sb_off.
template <int M>
__global__ void __launch_bounds__(256) transpose_partitions_kernel(const uint8_t* __restrict__ codes, int p0, int p1,
                                                                   const uint64_t* __restrict__ sb_off,      // [P] first superblock
                                                                   const uint64_t* __restrict__ vec_off,     // [P] first vector
                                                                   const uint32_t* __restrict__ size,        // [P]
                                                                   uint64_t sb_end,                          // first superblock after the run
                                                                   uint8_t* __restrict__ out) {              // native layout, whole database
    constexpr int CS = M / 2;
    const uint64_t sb_first = sb_off[p0];
    const uint64_t gidx = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint64_t sb = sb_first + gidx / 32;
    if (sb >= sb_end) return;
    const uint32_t lane = static_cast<uint32_t>(gidx % 32);
    int lo = p0, hi = p1;   // last p in [p0, p1) with sb_off[p] <= sb (empty partitions share their successor's offset)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (sb_off[mid] <= sb) lo = mid; else hi = mid;
    }
    const uint32_t n = size[lo];
    const uint32_t local = static_cast<uint32_t>(sb - sb_off[lo]) * kSbVec + lane * 8;
    const uint8_t* src = codes + (vec_off[lo] - vec_off[p0]) * CS;
    uint32_t word[M];
#pragma unroll
    for (int j = 0; j < M; ++j) word[j] = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t v = min(local + k, n - 1);
        const uint8_t* c = src + static_cast<size_t>(v) * CS;
#pragma unroll
        for (int b = 0; b < CS; ++b) {
            const uint32_t byte = c[b];
            word[2 * b] |= (byte & 15u) << (4 * k);
            word[2 * b + 1] |= (byte >> 4) << (4 * k);
        }
    }
    uint4* dst = reinterpret_cast<uint4*>(out + sb * sb_bytes(M));
#pragma unroll
    for (int q = 0; q < M / 4; ++q)
        dst[q * 32 + lane] = make_uint4(word[4 * q], word[4 * q + 1], word[4 * q + 2], word[4 * q + 3]);
}

// Inverse (tests: layout round trip).
template <int M>
__global__ void __launch_bounds__(256) untranspose_codes_kernel(const uint8_t* __restrict__ native, uint32_t size,
                                                                uint8_t* __restrict__ codes) {
    constexpr int CS = M / 2;
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= size) return;
    const uint32_t sb = v / kSbVec, lane = (v % kSbVec) / 8, k = v % 8;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(native + static_cast<size_t>(sb) * sb_bytes(M));
#pragma unroll
    for (int b = 0; b < CS; ++b) {
        const int j0 = 2 * b, j1 = 2 * b + 1;
        const uint32_t lo = (w[((j0 / 4) * 32 + lane) * 4 + (j0 % 4)] >> (4 * k)) & 15u;
        const uint32_t hi = (w[((j1 / 4) * 32 + lane) * 4 + (j1 % 4)] >> (4 * k)) & 15u;
        codes[static_cast<size_t>(v) * CS + b] = static_cast<uint8_t>(lo | (hi << 4));
    }
}

// ------------------------------------------------------------------------------------------
// Shared pieces of the scan kernels
// ------------------------------------------------------------------------------------------
// Append every vector of a group whose sum is below the bound.  pos0 = position of vector
// k=0 inside the (local) partition.
__device__ __forceinline__ void emit_candidates(const GroupAcc& g, uint32_t bound, uint32_t pos0, uint32_t size,
                                                uint32_t pos_base, uint32_t probe_rank, WarpList& list,
                                                int* hist = nullptr) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t v = lane_raw(g, k);
        if (!(v & 0x8000u) && pos0 + k < size) {
            const uint32_t d = v + bound - 0x8000u;
            list.push(make_key(d, probe_rank, pos_base + pos0 + k));
            if (hist) atomicAdd(hist + d, 1);
        }
    }
}

// CTA-wide candidate histogram (128 bins per query): every candidate ever pushed by any warp of the
// CTA is counted, so the smallest distance whose cumulative count reaches r is a valid value for
// the query's shared bound long before any single warp has collected r candidates of its own.
// Called by a whole warp; returns the bound (or 127 if fewer than r candidates were counted).
__device__ __forceinline__ int hist_bound(const int* hist, int r, int lane) {
    const int4 c = *reinterpret_cast<const int4*>(hist + lane * 4);
    const int mine = c.x + c.y + c.z + c.w;
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    const unsigned reached = __ballot_sync(0xffffffffu, incl >= r);
    if (!reached) return 127;
    const int src = __ffs(reached) - 1;
    int b = 127;
    if (lane == src) {
        const int c0 = incl - mine + c.x, c1 = c0 + c.y, c2 = c1 + c.z;
        b = lane * 4 + (c0 >= r ? 0 : (c1 >= r ? 1 : (c2 >= r ? 2 : 3)));
    }
    return __shfl_sync(0xffffffffu, b, src);
}

// One pair of sub-quantisers; the first pair of a vector starts the accumulators at -bound.
__device__ __forceinline__ void scan_pair(bool first, uint32_t w0, uint32_t w1, const uint4& t0, const uint4& t1,
                                          GroupAcc& g, const PipeK& k, uint32_t bound) {
    if (first) lut_pair<true>(w0, w1, t0, t1, g, k, 0u - bound);
    else lut_pair<false>(w0, w1, t0, t1, g, k, 0u);
}

// Writes a warp's final sorted list (r keys, kEmptyKey padded) to global memory.
__device__ __forceinline__ void store_list(const WarpList& list, uint64_t* dst, int r, int lane) {
    for (int i = lane; i < r; i += 32) dst[i] = list.keys[i];
}

// The query's shared bound: the smallest distance v for which some list already holds r records
// with d <= v (seeded from the keep-prefix by prefix_bound kernels).  A vector can only be in
// the top r if d <= v.  Read with a volatile load so every tile sees recent updates.
__device__ __forceinline__ int load_shared_bound(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

// Same, predicated on sel == 0 and written straight into `v` (a predicated load into the variable's own register:
// the C++ form goes through a temporary whose copy waits for the load at once).
__device__ __forceinline__ void load_shared_bound_if(int& v, const int* p, uint32_t sel) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.eq.u32 p, %2, 0;\n\t"
        "@p ld.volatile.global.s32 %0, [%1];\n\t}"
        : "+r"(v)
        : "l"(p), "r"(sel)
        : "memory");
}

__device__ __forceinline__ void load_shared_bound_now(int& v, const int* p) {
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
}

// After a warp pushed `added` candidates: count them CTA-wide and, when the count passes the next
// threshold, re-derive the query's shared bound from the CTA histogram.  Whole warp calls it;
// returns the new bound (or 127 when nothing changed).
// (global_too: the CTA shares a query with other CTAs — the threshold then starts at r / 4 instead of r and passing it
// returns -1: the caller publishes to the query's global histogram, hist_publish, which yields the bound.)
__device__ __forceinline__ int hist_update(int* hist, int* hist_total, int* hist_next, int added, int r, int lane,
                                           int* shared_bound, bool global_too = false) {
    int total = 0;
    if (lane == 0) {
        total = atomicAdd(hist_total, added) + added;
        // lane 0 alone decides (warp-uniform); the threshold is shared by the CTA's warps, so it is
        // read and raised with atomics (a stale value would only move the moment of the next update)
        if (total < atomicAdd(hist_next, 0)) total = -1;
    }
    total = __shfl_sync(0xffffffffu, total, 0);
    if (total < 0) return 127;
    if (global_too) {
        if (lane == 0) atomicMax(hist_next, total + max(r / 4, 8));
        return -1;
    }
    const int hb = hist_bound(hist, r, lane);
    if (lane == 0) {
        atomicMax(hist_next, total + max(r / 4, 8));
        if (hb < 127) atomicMin(shared_bound, hb);
    }
    return hb;
}

// The same across the CTAs of a query (flat scans cut a database into up to 148 x 16 chunks per query): the CTA histogram
// only knows the CTA's share of the vectors, so its r-th distance is that of 1/chunks of the database — on rank 0 of an
// 8-way sharded 1e9 scan the bound stayed 10-20 above the final one for most of a CTA's life, the pre-filter passed too
// many superblocks to the exact core and the scan took 2.78 ms where a perfect seed gave 2.565 (tools/exp_seed.py).
// When a query has several CTAs its candidates are therefore counted in a PENDING histogram (`pend`, the only one the
// CTA keeps then), and every r/4 candidates a warp moves what is pending — atomicExch per bin, so every candidate is
// handed over exactly once whichever warps publish concurrently — into the query's histogram in global memory (ghist,
// zeroed per launch: its counts never exceed the number of scanned vectors at each distance) and takes the bound of the
// union: the smallest distance whose global cumulative count reaches r.  A whole warp calls it; returns that bound or 127.
__device__ __forceinline__ int hist_publish(int* pend, int* ghist, int r, int lane, int* shared_bound) {
    int* p = pend + lane * 4;
    int* g = ghist + lane * 4;
    const int d0 = atomicExch(p, 0), d1 = atomicExch(p + 1, 0), d2 = atomicExch(p + 2, 0), d3 = atomicExch(p + 3, 0);
    if (d0) atomicAdd(g, d0);
    if (d1) atomicAdd(g + 1, d1);
    if (d2) atomicAdd(g + 2, d2);
    if (d3) atomicAdd(g + 3, d3);
    const int4 c = __ldcg(reinterpret_cast<const int4*>(g));   // the other CTAs' counts (own additions may still be on their way: fine)
    const int mine = c.x + c.y + c.z + c.w;
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    const unsigned reached = __ballot_sync(0xffffffffu, incl >= r);
    if (!reached) return 127;
    const int src = __ffs(reached) - 1;
    int b = 127;
    if (lane == src) {
        const int c0 = incl - mine + c.x, c1 = c0 + c.y, c2 = c1 + c.z;
        b = lane * 4 + (c0 >= r ? 0 : (c1 >= r ? 1 : (c2 >= r ? 2 : 3)));
        if (b < 127) atomicMin(shared_bound, b);
    }
    return __shfl_sync(0xffffffffu, b, src);
}

// ------------------------------------------------------------------------------------------
// Flat scan: grid = (chunks along the database, query groups of QB).  NW consumer warps +
// one producer warp per CTA; the producer streams tiles of NW superblocks through an NS-stage
// shared-memory ring with TMA bulk copies; consumer warp w owns superblock w of every tile.
// Each warp keeps one candidate list per query.
// ------------------------------------------------------------------------------------------
struct FlatScanArgs {
    const uint8_t* codes;   // native layout of the partition
    uint32_t n_sb;          // superblocks in the partition
    uint32_t size;          // vectors in the (local) partition
    uint32_t pos_base;      // position of local vector 0 in the full partition
    uint32_t sb_per_chunk;
    const int8_t* qtabs;    // [nq][M*16]
    int nq, r, cap;
    uint64_t* lists;        // [nq][n_lists][r]
    int n_lists;            // gridDim.x * NW
    int* shared_bound;      // [nq]
    PipeK k;                // {1, -1}
    int use_filter;         // register-table variant: clamped byte-lane pre-filter before the exact core
    int* ghist;             // [nq][128] the queries' global candidate histograms (zeroed per launch) or null
};

template <int M, int QB, int NW, int NS>
struct FlatCfg {
    static constexpr int kQuads = M / 4;
    static constexpr int kSbBytes = M * 128;
    static constexpr int kTileBytes = NW * kSbBytes;
    static constexpr bool kRegTab = (QB == 1 && M == 16 && NW <= 16);   // more warps: tables stay in shared memory
    static constexpr int kThreads = (NW + 1) * 32;
    static constexpr int kFiltBytes = kRegTab ? NW * M * 16 : 0;   // per-warp clamped table (pre-filter)
    // tiles | tables | filter tables | full/empty barriers | histogram + counters | list counts/bounds, rounded to 16 bytes
    static constexpr int kFixedBytes =
        ((NS * kTileBytes + QB * M * 16 + kFiltBytes + 2 * NS * 8 + QB * (128 + 2) * 4 + NW * QB * 8) + 15) / 16 * 16;
    static size_t smem_bytes(int cap) { return static_cast<size_t>(kFixedBytes) + static_cast<size_t>(NW) * QB * cap * 8; }
};

template <int M, int QB, int NW, int NS>
__global__ void __launch_bounds__((NW + 1) * 32, 1) scan_flat_kernel(const FlatScanArgs a) {
    using Cfg = FlatCfg<M, QB, NW, NS>;
    extern __shared__ __align__(128) uint8_t smem[];
    // every small array sits at a compile-time offset (no address registers live across the
    // lookup loop); only the candidate lists, whose size depends on `cap`, come last
    uint8_t* tiles = smem;
    uint4* qtab = reinterpret_cast<uint4*>(tiles + static_cast<size_t>(NS) * Cfg::kTileBytes);   // [QB][M]
    uint4* ftab = qtab + QB * M;                                                                  // [NW][M] (kRegTab only)
    uint64_t* full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(ftab) + Cfg::kFiltBytes);
    uint64_t* empty = full + NS;
    int* hist = reinterpret_cast<int*>(empty + NS);   // [QB][128] candidate histogram (16-byte aligned)
    int* hist_total = hist + QB * 128;                // [QB] candidates counted
    int* hist_next = hist_total + QB;                 // [QB] count at which the bound is recomputed
    int* cnt = hist_next + QB;                        // [NW][QB]
    int* bnd = cnt + NW * QB;
    uint64_t* lists = reinterpret_cast<uint64_t*>(tiles + Cfg::kFixedBytes);                     // [NW][QB][cap]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sb0 = min(blockIdx.x * a.sb_per_chunk, a.n_sb);
    const uint32_t sb1 = min(sb0 + a.sb_per_chunk, a.n_sb);
    const uint32_t n_tiles = (sb1 - sb0 + NW - 1) / NW;
    const int qbase = blockIdx.y * QB;
    const int nqb = min(QB, a.nq - qbase);
    const PipeK pk = a.k;

    // ---- init ----
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NW); }
        mbar_fence_init();
    }
    for (int i = threadIdx.x; i < QB * M; i += blockDim.x) {
        const int qi = i / M;
        qtab[i] = (qi < nqb) ? reinterpret_cast<const uint4*>(a.qtabs)[static_cast<size_t>(qbase) * M + i]
                             : make_uint4(0x7f7f7f7fu, 0x7f7f7f7fu, 0x7f7f7f7fu, 0x7f7f7f7fu);
    }
    for (int i = threadIdx.x; i < NW * QB * a.cap; i += blockDim.x) lists[i] = kEmptyKey;
    for (int i = threadIdx.x; i < NW * QB; i += blockDim.x) { cnt[i] = 0; bnd[i] = 127; }
    for (int i = threadIdx.x; i < QB * 128; i += blockDim.x) hist[i] = 0;
    for (int i = threadIdx.x; i < QB; i += blockDim.x) { hist_total[i] = 0; hist_next[i] = a.r; }
    __syncthreads();

    if (warp == NW) {
        // ---- producer: one elected lane drives the TMA ring ----
        if (lane == 0) {
            for (uint32_t t = 0; t < n_tiles; ++t) {
                const int s = t % NS;
                if (t >= NS) mbar_wait(&empty[s], ((t / NS) & 1) ^ 1);
                const uint32_t tsb = sb0 + t * NW;
                const uint32_t bytes = min(static_cast<uint32_t>(NW), sb1 - tsb) * Cfg::kSbBytes;
                mbar_arrive_expect_tx(&full[s], bytes);
                bulk_g2s(tiles + static_cast<size_t>(s) * Cfg::kTileBytes,
                         a.codes + static_cast<size_t>(tsb) * Cfg::kSbBytes, bytes, &full[s]);
            }
        }
#ifdef QADC_FINAL_SYNC
        __syncthreads();
#endif
        return;
    }

    // ---- consumers ----
    WarpList wl[QB];
#pragma unroll
    for (int qi = 0; qi < QB; ++qi) {
        wl[qi].keys = lists + (static_cast<size_t>(warp) * QB + qi) * a.cap;
        wl[qi].count = cnt + warp * QB + qi;
        wl[qi].bound = bnd + warp * QB + qi;
    }
    const int halves = (a.cap < a.r + kSbVec) ? 2 : 1;
    const int compact_at = min(a.cap - kSbVec / halves, 2 * a.r);

    // loop state kept in registers: this warp's superblock, its slot inside a stage, ring position
    uint32_t sb = sb0 + warp;
    uint32_t n_left = (sb < sb1) ? (sb1 - sb + NW - 1) / NW : 0;   // tiles in which this warp has a superblock
    uint32_t slot = smem_u32(tiles) + warp * Cfg::kSbBytes + lane * 16;
    uint32_t full_a = smem_u32(full), empty_a = smem_u32(empty);
    pin(slot); pin(sb); pin(n_left);   // full_a / empty_a are compile-time offsets from the shared-memory base
    uint32_t stage = 0, phase = 0;
    const uint32_t rt_zero = pk.one - 1u;   // 0 (PipeK::one is 1, a kernel argument the compiler cannot fold)
    int lbound[QB], gb[QB];   // strict local bound, shared bound (refreshed once per ring revolution)
#pragma unroll
    for (int qi = 0; qi < QB; ++qi) {
        lbound[qi] = 127;
        gb[qi] = (qi < nqb) ? load_shared_bound(a.shared_bound + qbase + qi) : 0;
    }

    // the rare path of one (superblock, query): append the candidates of the group sums `g`, count them
    // CTA-wide (histogram -> shared bound) and compact the list when it fills up
    auto emit_group = [&](const GroupAcc& g, const bool mine, const uint32_t bound, const int qi, const uint32_t sb) {
        // small lists (cap < r + 256) take the superblock in two position-ordered halves
        for (int half = 0; half < halves; ++half) {
            const int before = *wl[qi].count;
            __syncwarp();
            const bool my_turn = halves == 1 || (lane >> 4) == half;
            if (mine && my_turn)
                emit_candidates(g, bound, sb * kSbVec + lane * 8, a.size, a.pos_base, 0, wl[qi], hist + qi * 128);
            __syncwarp();
            const int now = *wl[qi].count;
            {
                // CTA-wide count of candidates; when it passes the next threshold this warp
                // re-derives the query's shared bound from the CTA histogram
                const int hb = hist_update(hist + qi * 128, hist_total + qi, hist_next + qi, now - before,
                                           a.r, lane, a.shared_bound + qbase + qi);
                if (hb < gb[qi]) gb[qi] = hb;
            }
            if (now >= compact_at) {
                wl[qi].compact(a.cap, a.r, lane, a.shared_bound + qbase + qi);
                lbound[qi] = *wl[qi].bound;
            }
        }
    };

    if constexpr (Cfg::kRegTab) {
        // ---- one query per pass, its table in 64 registers ----
        // treg holds either the exact table or (pre-filter on) the clamped one; the filter starts on
        // and a warp falls back to the exact core for a while when too many superblocks pass it
        // (loose bound, saturated tables), so the worst case costs what the exact core costs.
        uint4 treg[M];
        uint4* my_ftab = ftab + warp * M;
        // bound = the strict bound of the moment (min of the local strict bound and the shared bound + 1), the
        // only bound kept across iterations; f_start = start value of the filter lanes (127 - t_f in every byte,
        // t_f = the threshold the clamped table in treg was built for); ctl = mode control: filter on -> pass-rate
        // score (+6 per passing superblock, -1 otherwise, floor 0), filter off -> superblocks left before the
        // filter is tried again
        int bound = min(lbound[0], gb[0] + 1);
        uint32_t f_start = 0;
        bool filt_on = a.use_filter != 0;
        int ctl = 0;
        auto load_table = [&]() {   // all lanes; (re)builds treg for the current mode and bound
            if (filt_on) {
                const int t_f = bound - 1;
                f_start = filt_start(t_f);
                const uint32_t cap4 = static_cast<uint32_t>(filt_cap(t_f, M)) * 0x01010101u;
                __syncwarp();
                if (lane < M) {
                    const uint4 t = qtab[lane];
                    my_ftab[lane] = make_uint4(__vminu4(t.x, cap4), __vminu4(t.y, cap4), __vminu4(t.z, cap4), __vminu4(t.w, cap4));
                }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < M; ++j) treg[j] = my_ftab[j];
            } else {
#pragma unroll
                for (int j = 0; j < M; ++j) treg[j] = qtab[j];
            }
        };
        load_table();
        int gb_pending = bound - 1;

        uint32_t t = 0;
        for (; t < n_left; ++t) {
            mbar_wait_a(full_a + stage * 8, phase);
            const uint32_t src = slot + stage * Cfg::kTileBytes;
            uint4 w[Cfg::kQuads];
#pragma unroll
            for (int q = 0; q < Cfg::kQuads; ++q) w[q] = lds128(src + q * 512);
            const uint32_t landed = zero_after_loads(w[Cfg::kQuads - 1].w, rt_zero);   // 0, available once the loads have landed
            __syncwarp();
            if (lane == 0) mbar_arrive_a(empty_a + stage * 8 + landed);   // the words are in registers: release the stage early
            bool reload = false;
            if (stage == 0) {
                // the shared bound is read once per ring revolution and used one revolution later: the load
                // (an L2 hit, ~1 us under load) is never waited for
                bound = min(bound, gb_pending + 1);
                load_shared_bound_now(gb_pending, a.shared_bound + qbase);
                // rebuild the clamped table when the bound has moved enough to allow a tighter one
                reload = filt_on && bound + 3 <= 127 - static_cast<int>(f_start & 0xffu);
            }
            GroupAcc g;
            bool mine = false;
            if (filt_on) {
#ifdef QADC_FILT_ACC4
                // two accumulator pairs (even / odd quads): half the dependent-add chain length
                FiltAcc f{f_start, f_start}, f2{0u, 0u};
#pragma unroll
                for (int q = 0; q < Cfg::kQuads; ++q) {
                    FiltAcc& fq = (q & 1) ? f2 : f;
                    filt_word(w[q].x, treg[4 * q], fq, pk);
                    filt_word(w[q].y, treg[4 * q + 1], fq, pk);
                    filt_word(w[q].z, treg[4 * q + 2], fq, pk);
                    filt_word(w[q].w, treg[4 * q + 3], fq, pk);
                }
                f.a = madd(f2.a, f.a, pk); f.b = madd(f2.b, f.b, pk);
#else
                FiltAcc f{f_start, f_start};
#pragma unroll
                for (int q = 0; q < Cfg::kQuads; ++q) {
                    filt_word(w[q].x, treg[4 * q], f, pk);
                    filt_word(w[q].y, treg[4 * q + 1], f, pk);
                    filt_word(w[q].z, treg[4 * q + 2], f, pk);
                    filt_word(w[q].w, treg[4 * q + 3], f, pk);
                }
#endif
                ctl = max(ctl - 1, 0);
                if (__any_sync(0xffffffffu, filt_any(f))) {
                    // some vector may be below the bound: exact sums with the table read from shared memory
                    // (the words pass through an opaque IMAD so that ptxas does not keep the filter's 48
                    // selector registers alive, i.e. spilled, for reuse here)
                    ctl += 6;
                    if (ctl > 96) { filt_on = false; ctl = 256; reload = true; }   // ~1 superblock in 6 passes the filter
#pragma unroll
                    for (int p = 0; p < M / 2; ++p) {
                        const uint4 t0 = qtab[2 * p], t1 = qtab[2 * p + 1];
                        const uint4& wq = w[p >> 1];
                        scan_pair(p == 0, madd((p & 1) ? wq.z : wq.x, 0u, pk), madd((p & 1) ? wq.w : wq.y, 0u, pk), t0, t1, g, pk, bound);
                    }
                    mine = any_below(g);
                }
            } else {
#pragma unroll
                for (int p = 0; p < M / 2; ++p) {   // pairs of sub-quantisers
                    const uint4& wq = w[p >> 1];
                    scan_pair(p == 0, (p & 1) ? wq.z : wq.x, (p & 1) ? wq.w : wq.y, treg[2 * p], treg[2 * p + 1], g, pk, bound);
                }
                mine = any_below(g);
                if (--ctl <= 0 && a.use_filter) { filt_on = true; ctl = 0; reload = true; }
            }
            if (__any_sync(0xffffffffu, mine)) {   // rare: some vector of the superblock is a candidate
                lbound[0] = bound; gb[0] = bound - 1;   // equivalent split of the one strict bound
                emit_group(g, mine, bound, 0, sb0 + warp + t * NW);
                bound = min(bound, min(lbound[0], gb[0] + 1));
                reload = true;   // the 64 table registers need not stay live (or be spilled) across the rare path
            }
            if (reload) load_table();
            if (++stage == NS) { stage = 0; phase ^= 1; }
        }
        for (; t < n_tiles; ++t) {   // the chunk's last tile may hold no superblock for this warp
            mbar_wait_a(full_a + stage * 8, phase);
            __syncwarp();
            if (lane == 0) mbar_arrive_a(empty_a + stage * 8);
            if (++stage == NS) { stage = 0; phase ^= 1; }
        }
    } else {
        // ---- QB queries share every pass; their tables are read from shared memory per pair ----
        for (uint32_t t = 0; t < n_tiles; ++t) {
            mbar_wait_a(full_a + stage * 8, phase);
            if (t < n_left) {
                const uint32_t src = slot + stage * Cfg::kTileBytes;
                uint4 w[Cfg::kQuads];
#pragma unroll
                for (int q = 0; q < Cfg::kQuads; ++q) w[q] = lds128(src + q * 512);
                int gb_next[QB];
                if (stage == 0) {
#pragma unroll
                    for (int qi = 0; qi < QB; ++qi) gb_next[qi] = (qi < nqb) ? load_shared_bound(a.shared_bound + qbase + qi) : 0;
                }
                const uint32_t landed = zero_after_loads(w[Cfg::kQuads - 1].w, rt_zero);   // 0, available once the loads have landed
                __syncwarp();
                if (lane == 0) mbar_arrive_a(empty_a + stage * 8 + landed);   // the words are in registers: release the stage early
                // quad-outer, query-inner: the selectors of a quad's four code words (XOR + two shifts per word, on
                // the ALU pipe that binds this regime) are prepared once and used by all QB queries
                GroupAcc g[QB];
                uint32_t bound[QB];
#pragma unroll
                for (int qi = 0; qi < QB; ++qi) bound[qi] = static_cast<uint32_t>(min(lbound[qi], gb[qi] + 1));
                // (Early abandon — skipping the last sub-quantisers once every vector of the superblock has reached
                // the bound — is exact but was measured 5 % slower: the vote/branch splits the unrolled lookup chain.)
#pragma unroll
                for (int q = 0; q < Cfg::kQuads; ++q) {
                    const Sel s0 = make_sel(w[q].x, pk), s1 = make_sel(w[q].y, pk), s2 = make_sel(w[q].z, pk), s3 = make_sel(w[q].w, pk);
#pragma unroll
                    for (int qi = 0; qi < QB; ++qi) {
                        const uint4* tq = qtab + qi * M + 4 * q;
                        const uint4 t0 = tq[0], t1 = tq[1], t2 = tq[2], t3 = tq[3];
                        if (q == 0) lut_pair_sel<true>(s0, s1, t0, t1, g[qi], pk, 0u - bound[qi]);
                        else lut_pair_sel<false>(s0, s1, t0, t1, g[qi], pk, 0u);
                        lut_pair_sel<false>(s2, s3, t2, t3, g[qi], pk, 0u);
                    }
                }
#pragma unroll
                for (int qi = 0; qi < QB; ++qi) {
                    if (qi < nqb) {   // (a missing query of the last group has all-127 tables: it never has a candidate)
                        const bool mine = any_below(g[qi]);
                        if (__any_sync(0xffffffffu, mine)) emit_group(g[qi], mine, bound[qi], qi, sb);   // rare
                    }
                }
                if (stage == 0) {
#pragma unroll
                    for (int qi = 0; qi < QB; ++qi) gb[qi] = min(gb[qi], gb_next[qi]);
                }
            } else {
                __syncwarp();
                if (lane == 0) mbar_arrive_a(empty_a + stage * 8);
            }
            sb += NW;
            if (++stage == NS) { stage = 0; phase ^= 1; }
        }
    }

    // ---- final: one sorted list per (CTA, query) when the NW warp lists fit a CTA-wide sort in
    // the (now idle) tile ring, else one list per (warp, query) ----
    const bool cta_merge = a.n_lists == static_cast<int>(gridDim.x);
    uint64_t* scratch = reinterpret_cast<uint64_t*>(tiles);
    const int ctid = threadIdx.x;   // consumer thread index (consumer warps are 0..NW-1)
#pragma unroll
    for (int qi = 0; qi < QB; ++qi) {
        if (qi < nqb) wl[qi].compact(a.cap, a.r, lane, a.shared_bound + qbase + qi);
    }
    if (!cta_merge) {
#pragma unroll
        for (int qi = 0; qi < QB; ++qi)
            if (qi < nqb)
                store_list(wl[qi], a.lists + (static_cast<size_t>(qbase + qi) * a.n_lists + blockIdx.x * NW + warp) * a.r,
                           a.r, lane);
        return;
    }
    int n_sort = 64;
    while (n_sort < NW * a.r) n_sort <<= 1;
#pragma unroll
    for (int qi = 0; qi < QB; ++qi) {
        if (qi >= nqb) break;   // CTA-uniform
        named_bar_sync(1, NW * 32);   // scratch is free (all tiles consumed / previous query stored)
        for (int i = lane; i < a.r; i += 32) scratch[warp * a.r + i] = wl[qi].keys[i];
        for (int i = NW * a.r + ctid; i < n_sort; i += NW * 32) scratch[i] = kEmptyKey;
        named_bar_sync(1, NW * 32);
        bitonic_sort_u64(scratch, n_sort, ctid, NW * 32, NamedSync{1, NW * 32});
        uint64_t* dst = a.lists + (static_cast<size_t>(qbase + qi) * a.n_lists + blockIdx.x) * a.r;
        for (int i = ctid; i < a.r; i += NW * 32) dst[i] = scratch[i];
    }
}

// Final merge of a CTA's NW sorted warp lists into one sorted list of r keys (kEmptyKey-padded) in global memory.
// Only keys at or below the query's shared bound can be in its top r (at least r scanned vectors are at or below it), and
// the warp lists are sorted: each warp contributes that prefix of its list and the CTA sorts the next power of two above
// their total.  With one bound per query (hist_publish) that is a handful of keys per CTA instead of the 2048 slots of NW
// full lists — a 66-stage block-wide sort at the end of each of the up to 16 CTAs an SM runs per launch.  Every thread of
// the CTA calls it; `wcount` = NW ints of shared memory, `scratch` = max(r, total keys) u64 of shared memory that no warp
// uses any more.  Not inlined: its registers must not compete with the scan loop's (inlined, ptxas spilled a loop-carried
// address of the 128-register kernel).
template <int NW>
__device__ __noinline__ void cta_merge_bounded(const uint64_t* keys, uint64_t* scratch, int* wcount, const int* sbound,
                                               uint64_t* dst, int r) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();      // every warp has consumed its ring and left the candidate path
    const int vmax = load_shared_bound(sbound);
    int mine = 0;
    for (int base = 0; base < r; base += 32) {
        const int i = base + lane;
        const bool keep = i < r && keys[i] != kEmptyKey && static_cast<int>(keys[i] >> 48) <= vmax;
        mine += __popc(__ballot_sync(0xffffffffu, keep));
    }
    if (lane == 0) wcount[warp] = mine;
    __syncthreads();
    int off = 0, total = 0;
    for (int w2 = 0; w2 < NW; ++w2) { const int c = wcount[w2]; if (w2 < warp) off += c; total += c; }
    int n_sort = 64;
    while (n_sort < total) n_sort <<= 1;
    for (int i = lane; i < mine; i += 32) scratch[off + i] = keys[i];
    for (int i = total + threadIdx.x; i < max(n_sort, r); i += NW * 32) scratch[i] = kEmptyKey;
    __syncthreads();
    bitonic_sort_u64(scratch, n_sort, threadIdx.x, NW * 32, BlockSync());
    for (int i = threadIdx.x; i < r; i += NW * 32) dst[i] = scratch[i];
}

// ------------------------------------------------------------------------------------------
// Flat scan, one query per pass, 16x4 codes, PER-WARP rings: grid = (chunks, queries), NW warps per CTA, no producer
// warp.  Warp w owns superblocks sb0 + w, sb0 + w + NW, ... of the CTA's chunk (the CTA still sweeps its chunk
// front to back, so HBM sees 148 sequential streams) and feeds ITSELF: its elected lane keeps NSW one-superblock
// TMA bulk copies in flight into the warp's own shared-memory slots (one mbarrier per slot) and re-arms a slot as
// soon as the warp has copied that superblock into registers.  Compared with the CTA-wide ring of scan_flat_kernel
// (one producer warp, a stage is refilled only when ALL 15 consumer warps have released it) a slow warp no longer
// holds back the prefetch of the others — ncu showed 16 % of the warp time of that kernel waiting at the `full`
// barrier with HBM at 86 % of its measured peak — and the 16th warp computes instead of producing.
// The lookup / pre-filter / candidate machinery is the register-table path of scan_flat_kernel.
// ------------------------------------------------------------------------------------------
template <int NW, int NSW, int NPS = 1>
struct WarpRingCfg {
    static constexpr int M = 16, kQuads = 4, kSbBytes = 2048;
    static constexpr int kSlotBytes = NPS * kSbBytes;   // a ring slot holds NPS consecutive superblocks (one TMA copy, one barrier)
    static constexpr int kRingBytes = NW * NSW * kSlotBytes;
    static constexpr int kThreads = NW * 32;
    // rings | table | per-warp filter tables | barriers | histogram + counters | list counts/bounds, rounded to 16 bytes
    static constexpr int kFixedBytes = ((kRingBytes + M * 16 + NW * M * 16 + NW * NSW * 8 + (128 + 4) * 4 + NW * 8) + 15) / 16 * 16;
    static size_t smem_bytes(int cap) { return static_cast<size_t>(kFixedBytes) + static_cast<size_t>(NW) * cap * 8; }
};

template <int NW, int NSW, int NPS = 1>
__global__ void __launch_bounds__(NW * 32, 1) scan_flat_wr_kernel(const FlatScanArgs a) {
    using Cfg = WarpRingCfg<NW, NSW, NPS>;
    constexpr int M = 16;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* rings = smem;                                                                 // [NW][NSW][2048]
    uint4* qtab = reinterpret_cast<uint4*>(rings + Cfg::kRingBytes);                       // [M]
    uint4* ftab = qtab + M;                                                                // [NW][M]
    uint64_t* full = reinterpret_cast<uint64_t*>(ftab + NW * M);                           // [NW][NSW]
    int* hist = reinterpret_cast<int*>(full + NW * NSW);                                   // [128]
    int* hist_total = hist + 128;
    int* hist_next = hist_total + 1;
    int* cnt = hist_next + 3;                                                              // [NW]
    int* bnd = cnt + NW;
    uint64_t* lists = reinterpret_cast<uint64_t*>(smem + Cfg::kFixedBytes);                // [NW][cap]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sb0 = min(blockIdx.x * a.sb_per_chunk, a.n_sb);
    const uint32_t sb1 = min(sb0 + a.sb_per_chunk, a.n_sb);
    const int q = blockIdx.y;
    const PipeK pk = a.k;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NW * NSW; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    for (int i = threadIdx.x; i < M; i += blockDim.x) qtab[i] = reinterpret_cast<const uint4*>(a.qtabs)[static_cast<size_t>(q) * M + i];
    for (int i = threadIdx.x; i < NW * a.cap; i += blockDim.x) lists[i] = kEmptyKey;
    for (int i = threadIdx.x; i < NW; i += blockDim.x) { cnt[i] = 0; bnd[i] = 127; }
    int* ghist = (a.ghist && gridDim.x > 1) ? a.ghist + static_cast<size_t>(q) * 128 : nullptr;   // one chunk: the CTA sees everything
    for (int i = threadIdx.x; i < 128; i += blockDim.x) hist[i] = 0;   // (the pending counts when the query has several CTAs)
    if (threadIdx.x == 0) { *hist_total = 0; *hist_next = ghist ? max(a.r / 4, 8) : a.r; }
    __syncthreads();

    WarpList wl{lists + static_cast<size_t>(warp) * a.cap, cnt + warp, bnd + warp};
    const int halves = (a.cap < a.r + kSbVec) ? 2 : 1;
    const int compact_at = min(a.cap - kSbVec / halves, 2 * a.r);
    int* sbound = a.shared_bound + q;

    // this warp's units (NPS consecutive superblocks each): unit warp + i * NW of the chunk, i < n_mine; only the last
    // unit of the database can be short
    const uint32_t n_units = (sb1 - sb0 + NPS - 1) / NPS;
    const uint32_t first = sb0 + warp * NPS;
    const uint32_t n_mine = (static_cast<uint32_t>(warp) < n_units) ? (n_units - warp + NW - 1) / NW : 0;
    const uint32_t ring_a = smem_u32(rings) + warp * (NSW * Cfg::kSlotBytes);
    const uint32_t full_a = smem_u32(full) + warp * (NSW * 8);
    const uint8_t* src0 = a.codes + static_cast<size_t>(first) * Cfg::kSbBytes;
    auto unit_bytes = [&](uint32_t i) -> uint32_t {   // bytes of this warp's unit i
        if (NPS == 1) return Cfg::kSbBytes;
        const uint32_t sb = first + i * (NW * NPS);
        return min(static_cast<uint32_t>(NPS), sb1 - sb) * Cfg::kSbBytes;
    };
    auto issue = [&](uint32_t i, uint32_t slot) {   // elected lane: arm the slot's barrier and start the copy of unit i
        const uint32_t bytes = unit_bytes(i);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_a + slot * 8), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         ring_a + slot * Cfg::kSlotBytes),
                     "l"(src0 + static_cast<size_t>(i) * (NW * Cfg::kSlotBytes)), "r"(bytes), "r"(full_a + slot * 8)
                     : "memory");
    };
    if (lane == 0)
        for (uint32_t i = 0; i < min(n_mine, static_cast<uint32_t>(NSW)); ++i) issue(i, i);

    uint4 treg[M];
    uint4* my_ftab = ftab + warp * M;
    int lb = 127, gb = load_shared_bound(sbound);
    int bound = min(lb, gb + 1);
    uint32_t f_start = 0;
    bool filt_on = a.use_filter != 0;
    int ctl = 0;
    auto load_table = [&]() {
        if (filt_on) {
            const int t_f = bound - 1;
            f_start = filt_start(t_f);
            const uint32_t cap4 = static_cast<uint32_t>(filt_cap(t_f, M)) * 0x01010101u;
            __syncwarp();
            if (lane < M) {
                const uint4 t = qtab[lane];
                my_ftab[lane] = make_uint4(__vminu4(t.x, cap4), __vminu4(t.y, cap4), __vminu4(t.z, cap4), __vminu4(t.w, cap4));
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < M; ++j) treg[j] = my_ftab[j];
        } else {
#pragma unroll
            for (int j = 0; j < M; ++j) treg[j] = qtab[j];
        }
    };
    load_table();
    int gb_pending = bound - 1;
    uint32_t phase = 0;
    uint32_t ring_r = ring_a + lane * 16, ring_w = ring_a, full_r = full_a;   // kept in registers: the compiler otherwise
    pin(ring_r); pin(ring_w); pin(full_r);                                     // rebuilds them from the shared-window base
    const uint8_t* refill = src0 + static_cast<size_t>(NSW) * (NW * Cfg::kSlotBytes);   // source of unit i + NSW
    const uint32_t n_refill = n_mine > NSW ? n_mine - NSW : 0;                            // units that have a successor to prefetch

    auto body = [&](const uint32_t slot, const uint32_t i) {   // one unit = NPS superblocks behind one barrier
        mbar_wait_a(full_r + slot * 8, phase);
        bool reload = false;
        const uint32_t nsb = (NPS == 1) ? 1u : min(static_cast<uint32_t>(NPS), sb1 - (first + i * (NW * NPS)));
#pragma unroll 1
        for (uint32_t h = 0; h < nsb; ++h) {
            const uint32_t src = ring_r + slot * Cfg::kSlotBytes + h * Cfg::kSbBytes;
            uint4 w[4];
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) w[qd] = lds128(src + qd * 512);
            // rare: some vector of the superblock is a candidate
            auto emit = [&](const GroupAcc& g, const bool mine) {
                const uint32_t sb = first + i * (NW * NPS) + h;
                lb = bound; gb = bound - 1;
                for (int half = 0; half < halves; ++half) {
                    const int before = *wl.count;
                    __syncwarp();
                    const bool my_turn = halves == 1 || (lane >> 4) == half;
                    if (mine && my_turn) emit_candidates(g, bound, sb * kSbVec + lane * 8, a.size, a.pos_base, 0, wl, hist);
                    __syncwarp();
                    const int now = *wl.count;
                    int hb = hist_update(hist, hist_total, hist_next, now - before, a.r, lane, sbound, ghist != nullptr);
                    if (hb == -1) hb = hist_publish(hist, ghist, a.r, lane, sbound);   // threshold passed: share
                    if (hb < gb) gb = hb;
                    if (now >= compact_at) {
                        wl.compact(a.cap, a.r, lane, sbound);
                        lb = *wl.bound;
                    }
                }
                bound = min(bound, min(lb, gb + 1));
                reload = true;
            };
            if (filt_on) {
                FiltAcc f{f_start, f_start};
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                    filt_word(w[qd].x, treg[4 * qd], f, pk);
                    filt_word(w[qd].y, treg[4 * qd + 1], f, pk);
                    filt_word(w[qd].z, treg[4 * qd + 2], f, pk);
                    filt_word(w[qd].w, treg[4 * qd + 3], f, pk);
                }
                if (__any_sync(0xffffffffu, filt_any(f))) {
                    // some vector may be below the bound: exact sums, table read from shared memory (the words pass through an
                    // opaque IMAD so that ptxas does not keep the filter's selector registers alive for reuse here)
                    ctl += 6;
                    if (ctl > 96) { filt_on = false; ctl = 256; reload = true; }   // ~1 superblock in 6 passes the filter
                    GroupAcc g;
#pragma unroll
                    for (int p = 0; p < M / 2; ++p) {
                        const uint4 t0 = qtab[2 * p], t1 = qtab[2 * p + 1];
                        const uint4& wq = w[p >> 1];
                        scan_pair(p == 0, madd((p & 1) ? wq.z : wq.x, 0u, pk), madd((p & 1) ? wq.w : wq.y, 0u, pk), t0, t1, g, pk, bound);
                    }
                    const bool mine = any_below(g);
                    if (__any_sync(0xffffffffu, mine)) emit(g, mine);
                }
            } else {
                GroupAcc g;
#pragma unroll
                for (int p = 0; p < M / 2; ++p) {
                    const uint4& wq = w[p >> 1];
                    scan_pair(p == 0, (p & 1) ? wq.z : wq.x, (p & 1) ? wq.w : wq.y, treg[2 * p], treg[2 * p + 1], g, pk, bound);
                }
                const bool mine = any_below(g);
                if (__any_sync(0xffffffffu, mine)) emit(g, mine);
            }
            // The slot is refilled HERE, behind the vote that consumed every code word of the superblock: instructions
            // issue in order, so all four shared-memory loads have landed by now.  (Issued right behind the loads, as the
            // first version did, the copy depends on nothing they produce; a load still queued in the LSU when an
            // L2-hit TMA copy lands would then read the NEXT superblock — observed in the batched kernel, whose table
            // loads keep the LSU busy.  Costs nothing: three slots are in flight while this one is computed.)
            if (NPS == 1 || h + 1 == nsb) {
                __syncwarp();
                if (i < n_refill && elect_one()) {
                    const uint32_t bytes = unit_bytes(i + NSW);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_r + slot * 8), "r"(bytes) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     ring_w + slot * Cfg::kSlotBytes),
                                 "l"(refill), "r"(bytes), "r"(full_r + slot * 8)
                                 : "memory");
                }
            }
            if (slot == 0 && (NPS == 1 || h == 0)) {
                // once per ring revolution: the shared bound read one revolution ago (never waited for), the filter's
                // pass-rate score decays, a tighter clamped table is built when the bound has moved enough
                bound = min(bound, gb_pending + 1);
                load_shared_bound_now(gb_pending, sbound);
                ctl = filt_on ? max(ctl - NSW * NPS, 0) : ctl - NSW * NPS;
                if (filt_on) reload |= bound + 3 <= 127 - static_cast<int>(f_start & 0xffu);
                else if (ctl <= 0 && a.use_filter) { filt_on = true; ctl = 0; reload = true; }
            }
            // before the next superblock: a filter that was switched on or off needs its own tables
            if (reload) { load_table(); reload = false; }
        }
        refill += NW * Cfg::kSlotBytes;
    };

    // (Unrolling the NSW slots so that every address is base + immediate was measured 20 % SLOWER: the four copies of
    // the body no longer fit the instruction cache — `no_instruction` stalls 0.75 per issue.)
    uint32_t slot = 0;
    for (uint32_t i = 0; i < n_mine; ++i) {
        body(slot, i);
        if (++slot == NSW) { slot = 0; phase ^= 1; }
    }

    // ---- final: one sorted list per CTA (the NW warp lists merged in the idle ring) or one per warp ----
    wl.compact(a.cap, a.r, lane, sbound);
    const bool cta_merge = a.n_lists == static_cast<int>(gridDim.x);
    if (!cta_merge) {
        store_list(wl, a.lists + (static_cast<size_t>(q) * a.n_lists + blockIdx.x * NW + warp) * a.r, a.r, lane);
        return;
    }
    cta_merge_bounded<NW>(wl.keys, reinterpret_cast<uint64_t*>(rings), hist, sbound,
                          a.lists + (static_cast<size_t>(q) * a.n_lists + blockIdx.x) * a.r, a.r);
}

// ------------------------------------------------------------------------------------------
// Flat scan, QB queries per pass (any M), PER-WARP rings: the batched regime of configs 1 and 3 (many queries over a
// database that the L2 holds or that every pass streams).  Same ring as scan_flat_wr_kernel — warp w owns superblocks
// sb0 + w, sb0 + w + NW, ... and keeps NSW one-superblock TMA bulk copies in flight into its own slots — around the
// shared-selector lookup loop of scan_flat_kernel's batched path (tables of the QB queries in shared memory, selectors
// of a quad prepared once for all of them).  ncu on config 1 with the CTA-wide ring (profiles/r02_flat_batched_*):
// 9 % of all executed instructions were barrier spins (6 % of them the producer warp's), 10 % of the samples sat at the
// `full` barrier: a warp that takes the rare path (candidate lists, compaction) holds back the refill of the stage for
// the other 14.  Here nobody waits for anybody else and the 16th warp computes.
// ------------------------------------------------------------------------------------------
template <int M, int QB, int NW, int NSW>
struct WarpRingBatchCfg {
    static constexpr int kQuads = M / 4, kSbBytes = M * 128;
    static constexpr int kRingBytes = NW * NSW * kSbBytes;
    static constexpr int kThreads = NW * 32;
    // rings | tables | barriers | histograms + counters | list counts/bounds, rounded to 16 bytes
    static constexpr int kFixedBytes = ((kRingBytes + QB * M * 16 + NW * NSW * 8 + QB * (128 + 2) * 4 + NW * QB * 8) + 15) / 16 * 16;
    static size_t smem_bytes(int cap) { return static_cast<size_t>(kFixedBytes) + static_cast<size_t>(NW) * QB * cap * 8; }
};

template <int M, int QB, int NW, int NSW>
__global__ void __launch_bounds__(NW * 32, 1) scan_flat_wrq_kernel(const FlatScanArgs a) {
    using Cfg = WarpRingBatchCfg<M, QB, NW, NSW>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* rings = smem;                                                                 // [NW][NSW][kSbBytes]
    uint4* qtab = reinterpret_cast<uint4*>(rings + Cfg::kRingBytes);                       // [QB][M]
    uint64_t* full = reinterpret_cast<uint64_t*>(qtab + QB * M);                           // [NW][NSW]
    int* hist = reinterpret_cast<int*>(full + NW * NSW);                                   // [QB][128]
    int* hist_total = hist + QB * 128;                                                     // [QB]
    int* hist_next = hist_total + QB;                                                      // [QB]
    int* cnt = hist_next + QB;                                                             // [NW][QB]
    int* bnd = cnt + NW * QB;
    uint64_t* lists = reinterpret_cast<uint64_t*>(smem + Cfg::kFixedBytes);                // [NW][QB][cap]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sb0 = min(blockIdx.x * a.sb_per_chunk, a.n_sb);
    const uint32_t sb1 = min(sb0 + a.sb_per_chunk, a.n_sb);
    const int qbase = blockIdx.y * QB;
    const int nqb = min(QB, a.nq - qbase);
    const PipeK pk = a.k;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NW * NSW; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    for (int i = threadIdx.x; i < QB * M; i += blockDim.x) {
        const int qi = i / M;
        qtab[i] = (qi < nqb) ? reinterpret_cast<const uint4*>(a.qtabs)[static_cast<size_t>(qbase) * M + i]
                             : make_uint4(0x7f7f7f7fu, 0x7f7f7f7fu, 0x7f7f7f7fu, 0x7f7f7f7fu);
    }
    for (int i = threadIdx.x; i < NW * QB * a.cap; i += blockDim.x) lists[i] = kEmptyKey;
    for (int i = threadIdx.x; i < NW * QB; i += blockDim.x) { cnt[i] = 0; bnd[i] = 127; }
    int* const ghist0 = (a.ghist && gridDim.x > 1) ? a.ghist + static_cast<size_t>(qbase) * 128 : nullptr;   // one chunk: nothing to share
    for (int i = threadIdx.x; i < QB * 128; i += blockDim.x) hist[i] = 0;   // (the pending counts when the queries have several CTAs)
    for (int i = threadIdx.x; i < QB; i += blockDim.x) { hist_total[i] = 0; hist_next[i] = ghist0 ? max(a.r / 4, 8) : a.r; }
    __syncthreads();

    // (warp, query) candidate lists; the rare path takes the query as a RUNTIME index so that it exists once in the
    // code (inlined once per query it made the QB = 4 loop miss the instruction cache: no_instruction 3.4 stalls/issue)
    auto list_of = [&](int qi) { return WarpList{lists + (static_cast<size_t>(warp) * QB + qi) * a.cap, cnt + warp * QB + qi, bnd + warp * QB + qi}; };
    const int halves = (a.cap < a.r + kSbVec) ? 2 : 1;
    const int compact_at = min(a.cap - kSbVec / halves, 2 * a.r);

    // this warp's superblocks: sb0 + warp + i * NW, i < n_mine
    const uint32_t first = sb0 + warp;
    const uint32_t n_mine = (first < sb1) ? (sb1 - first + NW - 1) / NW : 0;
    const uint32_t ring_a = smem_u32(rings) + warp * (NSW * Cfg::kSbBytes);
    const uint32_t full_a = smem_u32(full) + warp * (NSW * 8);
    const uint8_t* src0 = a.codes + static_cast<size_t>(first) * Cfg::kSbBytes;
    if (lane == 0)
        for (uint32_t i = 0; i < min(n_mine, static_cast<uint32_t>(NSW)); ++i) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_a + i * 8), "r"(Cfg::kSbBytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             ring_a + i * Cfg::kSbBytes),
                         "l"(src0 + static_cast<size_t>(i) * (NW * Cfg::kSbBytes)), "r"(Cfg::kSbBytes), "r"(full_a + i * 8)
                         : "memory");
        }

    int lbound[QB], gb[QB];   // strict local bound, shared bound (refreshed once per ring revolution)
#pragma unroll
    for (int qi = 0; qi < QB; ++qi) {
        lbound[qi] = 127;
        gb[qi] = (qi < nqb) ? load_shared_bound(a.shared_bound + qbase + qi) : 0;
    }
    // rare path of one (superblock, query): returns the shared-bound candidate in hb and the list's strict bound in lb
    auto emit_group = [&](const GroupAcc& g, const bool mine, const uint32_t bound, const int qi, const uint32_t sb, int& hb_out, int& lb_out) {
        WarpList wl = list_of(qi);
        for (int half = 0; half < halves; ++half) {
            const int before = *wl.count;
            __syncwarp();
            const bool my_turn = halves == 1 || (lane >> 4) == half;
            if (mine && my_turn) emit_candidates(g, bound, sb * kSbVec + lane * 8, a.size, a.pos_base, 0, wl, hist + qi * 128);
            __syncwarp();
            const int now = *wl.count;
            int hb = hist_update(hist + qi * 128, hist_total + qi, hist_next + qi, now - before, a.r, lane,
                                 a.shared_bound + qbase + qi, ghist0 != nullptr);
            if (hb == -1)   // threshold passed: pool the candidates of the query's chunks (hist_publish)
                hb = hist_publish(hist + qi * 128, ghist0 + qi * 128, a.r, lane, a.shared_bound + qbase + qi);
            if (hb < hb_out) hb_out = hb;
            if (now >= compact_at) {
                wl.compact(a.cap, a.r, lane, a.shared_bound + qbase + qi);
                lb_out = *wl.bound;
            }
        }
    };

    uint32_t phase = 0;
    uint32_t ring_r = ring_a + lane * 16, ring_w = ring_a, full_r = full_a;
    pin(ring_r); pin(ring_w); pin(full_r);
    const uint8_t* refill = src0 + static_cast<size_t>(NSW) * (NW * Cfg::kSbBytes);   // source of superblock i + NSW
    const uint32_t n_refill = n_mine > NSW ? n_mine - NSW : 0;
    uint32_t slot = 0;
    for (uint32_t i = 0; i < n_mine; ++i) {
        mbar_wait_a(full_r + slot * 8, phase);
        const uint32_t src = ring_r + slot * Cfg::kSbBytes;
        uint4 w[Cfg::kQuads];
#pragma unroll
        for (int q = 0; q < Cfg::kQuads; ++q) w[q] = lds128(src + q * 512);
        int gb_next[QB];
        if (slot == 0) {
#pragma unroll
            for (int qi = 0; qi < QB; ++qi) gb_next[qi] = (qi < nqb) ? load_shared_bound(a.shared_bound + qbase + qi) : 0;
        }
        GroupAcc g[QB];
        uint32_t bound[QB];
#pragma unroll
        for (int qi = 0; qi < QB; ++qi) bound[qi] = static_cast<uint32_t>(min(lbound[qi], gb[qi] + 1));
#pragma unroll
        for (int q = 0; q < Cfg::kQuads; ++q) {
            const Sel s0 = make_sel(w[q].x, pk), s1 = make_sel(w[q].y, pk), s2 = make_sel(w[q].z, pk), s3 = make_sel(w[q].w, pk);
#pragma unroll
            for (int qi = 0; qi < QB; ++qi) {
                const uint4* tq = qtab + qi * M + 4 * q;
                const uint4 t0 = tq[0], t1 = tq[1], t2 = tq[2], t3 = tq[3];
                if (q == 0) lut_pair_sel<true>(s0, s1, t0, t1, g[qi], pk, 0u - bound[qi]);
                else lut_pair_sel<false>(s0, s1, t0, t1, g[qi], pk, 0u);
                lut_pair_sel<false>(s2, s3, t2, t3, g[qi], pk, 0u);
            }
        }
        // one vote for all QB queries (a missing query of the last group has all-127 tables: it never has a candidate)
        uint32_t mine_mask = 0;
#pragma unroll
        for (int qi = 0; qi < QB; ++qi) mine_mask |= any_below(g[qi]) ? (1u << qi) : 0u;
        if (__any_sync(0xffffffffu, mine_mask != 0)) {   // rare: some vector of the superblock is a candidate of some query
            uint32_t pend = __reduce_or_sync(0xffffffffu, mine_mask);
#pragma unroll 1
            while (pend) {
                const int qi = __ffs(pend) - 1;
                pend &= pend - 1;
                GroupAcc gs = g[0];
                uint32_t bs = bound[0];
#pragma unroll
                for (int u = 1; u < QB; ++u)
                    if (u == qi) { gs = g[u]; bs = bound[u]; }
                int hb = 127, lb = 127;
                emit_group(gs, (mine_mask >> qi) & 1u, bs, qi, first + i * NW, hb, lb);
#pragma unroll
                for (int u = 0; u < QB; ++u)
                    if (u == qi) { gb[u] = min(gb[u], hb); lbound[u] = min(lbound[u], lb); }
            }
        }
        // The slot is refilled behind the vote that consumed every code word of the superblock: instructions issue in
        // order, so the shared-memory loads of the slot have landed.  (Issued right behind the loads, the copy depends on
        // nothing they produce, and the table loads of the other warps can hold a code-word load in the LSU queue longer
        // than an L2-hit TMA copy takes: the warp then read words of the NEXT superblock — caught by the variant-agreement
        // test, ~3 % of the queries of a 4-queries-per-pass run.)
        __syncwarp();
        if (i < n_refill && elect_one()) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_r + slot * 8), "r"(Cfg::kSbBytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             ring_w + slot * Cfg::kSbBytes),
                         "l"(refill), "r"(Cfg::kSbBytes), "r"(full_r + slot * 8)
                         : "memory");
        }
        refill += NW * Cfg::kSbBytes;
        if (slot == 0) {
#pragma unroll
            for (int qi = 0; qi < QB; ++qi) gb[qi] = min(gb[qi], gb_next[qi]);
        }
        if (++slot == NSW) { slot = 0; phase ^= 1; }
    }

    // ---- final: one sorted list per (CTA, query) when the NW warp lists fit a CTA-wide sort in the idle rings,
    // else one list per (warp, query) ----
    const bool cta_merge = a.n_lists == static_cast<int>(gridDim.x);
    uint64_t* scratch = reinterpret_cast<uint64_t*>(rings);
#pragma unroll 1
    for (int qi = 0; qi < nqb; ++qi)
        list_of(qi).compact(a.cap, a.r, lane, a.shared_bound + qbase + qi);
    if (!cta_merge) {
#pragma unroll 1
        for (int qi = 0; qi < nqb; ++qi)
                store_list(list_of(qi), a.lists + (static_cast<size_t>(qbase + qi) * a.n_lists + blockIdx.x * NW + warp) * a.r, a.r, lane);
        return;
    }
    // Only keys at or below the query's shared bound can be in its top r (at least r scanned vectors are at or below it),
    // and the warp lists are sorted: each warp contributes that prefix of its list, and the CTA sorts the next power of
    // two above their total (typically 128-256 keys instead of the 2048 slots of NW full lists: the full sort was 6 % of
    // the kernel's instructions on config 1).
    int* wcount = hist;   // [NW], the histograms are no longer needed
#pragma unroll 1
    for (int qi = 0; qi < nqb; ++qi) {   // CTA-uniform
        __syncthreads();        // every warp has consumed its ring / the previous query is stored
        const int vmax = load_shared_bound(a.shared_bound + qbase + qi);
        const uint64_t* keys = list_of(qi).keys;
        int mine = 0;
        for (int base = 0; base < a.r; base += 32) {
            const int i = base + lane;
            const bool keep = i < a.r && keys[i] != kEmptyKey && static_cast<int>(keys[i] >> 48) <= vmax;
            mine += __popc(__ballot_sync(0xffffffffu, keep));
        }
        if (lane == 0) wcount[warp] = mine;
        __syncthreads();
        int off = 0, total = 0;
        for (int w2 = 0; w2 < NW; ++w2) { const int c = wcount[w2]; if (w2 < warp) off += c; total += c; }
        int n_sort = 64;
        while (n_sort < total) n_sort <<= 1;
        for (int i = lane; i < mine; i += 32) scratch[off + i] = keys[i];
        for (int i = total + threadIdx.x; i < max(n_sort, a.r); i += NW * 32) scratch[i] = kEmptyKey;
        __syncthreads();
        bitonic_sort_u64(scratch, n_sort, threadIdx.x, NW * 32, BlockSync());
        uint64_t* dst = a.lists + (static_cast<size_t>(qbase + qi) * a.n_lists + blockIdx.x) * a.r;
        for (int i = threadIdx.x; i < a.r; i += NW * 32) dst[i] = scratch[i];
    }
}

// ------------------------------------------------------------------------------------------
// IVF scan: grid = (probe chunks, queries).  The CTA cuts the inverted lists of its probes into
// work items of `sb_per_item` superblocks, numbered in canonical order (probe rank, position),
// and its warps claim them from a shared counter: a warp therefore still visits its vectors in
// increasing canonical order (what the strict pass rule of WarpList needs) while long and short
// lists, and probes that are empty on this shard, no longer unbalance the warps.  Per item a
// warp copies that probe's int8 table into its shared-memory slot (kept while consecutive items
// belong to the same probe) and streams the superblocks with coalesced 128-bit loads; with the
// tables out of the register file 3-4 CTAs fit an SM and the load latency is covered by
// occupancy.  (A per-warp TMA ring with one-item lookahead was measured slower.)
// ------------------------------------------------------------------------------------------
struct IvfScanArgs {
    const uint8_t* codes;            // native layout, all partitions
    const uint64_t* part_sb_off;     // [K] first superblock of partition p
    const uint32_t* part_size;       // [K]
    const uint32_t* part_pos_base;   // [K]
    const int32_t* assign;           // [nq][ma]
    const int8_t* qtabs;             // [nq][ma][M*16]
    int nq, ma, r, cap, probes_per_chunk, sb_per_item;
    uint64_t* lists;                 // [nq][n_lists][r]
    int n_lists;                     // gridDim.x * NW
    int* shared_bound;               // [nq]
    PipeK k;
};

// dynamic shared memory of scan_ivf_kernel: tables | lists | counts, bounds | histogram | probe metadata
__host__ __device__ inline size_t ivf_smem_bytes(int m, int nw, int cap, int probes_per_chunk) {
    size_t b = static_cast<size_t>(nw) * m * 16 + static_cast<size_t>(nw) * cap * 8 + static_cast<size_t>(nw) * 8;
    b = (b + 15) / 16 * 16 + (128 + 4) * 4;                      // hist, hist_total, hist_next, next_item, pad
    b += static_cast<size_t>(probes_per_chunk) * (8 + 4 + 4) + (static_cast<size_t>(probes_per_chunk) + 1) * 4;
    return b;
}

template <int M, int NW>
__global__ void __launch_bounds__(NW * 32, M == 16 ? 4 : 3) scan_ivf_kernel(const IvfScanArgs a) {
    constexpr int kQuads = M / 4, kSbBytes = M * 128;
    extern __shared__ __align__(128) uint8_t smem[];
    uint4* wtab = reinterpret_cast<uint4*>(smem) + (threadIdx.x >> 5) * M;       // this warp's int8 table [M]
    uint64_t* lists = reinterpret_cast<uint64_t*>(smem + NW * M * 16);           // [NW][cap]
    int* cnt = reinterpret_cast<int*>(lists + static_cast<size_t>(NW) * a.cap);  // [NW]
    int* bnd = cnt + NW;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.y;
    const int a0 = blockIdx.x * a.probes_per_chunk, a1 = min(a0 + a.probes_per_chunk, a.ma);
    const int n_probe = a1 - a0;
    const PipeK pk = a.k;

    int* hist = reinterpret_cast<int*>((reinterpret_cast<uintptr_t>(bnd + NW) + 15) & ~uintptr_t(15));   // [128]
    int* hist_total = hist + 128;
    int* hist_next = hist_total + 1;
    int* next_item = hist_next + 1;
    uint64_t* m_off = reinterpret_cast<uint64_t*>(hist + 128 + 4);               // [ppc] first superblock of the probe's list
    uint32_t* m_size = reinterpret_cast<uint32_t*>(m_off + a.probes_per_chunk);  // [ppc]
    uint32_t* m_pos = m_size + a.probes_per_chunk;                               // [ppc]
    int* item_off = reinterpret_cast<int*>(m_pos + a.probes_per_chunk);          // [ppc + 1] first item of probe i

    for (int i = threadIdx.x; i < 128; i += blockDim.x) hist[i] = 0;
    if (threadIdx.x == 0) { *hist_total = 0; *hist_next = a.r; *next_item = 0; }
    WarpList wl{lists + static_cast<size_t>(warp) * a.cap, cnt + warp, bnd + warp};
    for (int i = lane; i < a.cap; i += 32) wl.keys[i] = kEmptyKey;
    if (lane == 0) { *wl.count = 0; *wl.bound = 127; }
    // probe metadata and the item numbering (warp 0: chunked warp scan over the probes)
    if (warp == 0) {
        int running = 0;
        for (int base = 0; base < n_probe; base += 32) {
            const int i = base + lane;
            int items = 0;
            if (i < n_probe) {
                const int p = a.assign[static_cast<size_t>(q) * a.ma + a0 + i];
                const uint32_t size = a.part_size[p];   // 0: nothing to scan (db_query_4.cpp:291-293)
                m_size[i] = size;
                m_pos[i] = a.part_pos_base[p];
                m_off[i] = a.part_sb_off[p];
                const uint32_t n_sb = (size + kSbVec - 1) / kSbVec;
                items = static_cast<int>((n_sb + a.sb_per_item - 1) / a.sb_per_item);
            }
            int incl = items;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (i < n_probe) item_off[i + 1] = running + incl;
            running += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) item_off[0] = 0;
    }
    __syncthreads();
    const int total_items = item_off[n_probe];
    const int compact_at = min(a.cap - kSbVec, 2 * a.r);
    int* sbound = a.shared_bound + q;

    int cur = -1;   // probe (index in the chunk) whose table is in wtab
    for (;;) {
        int it = 0;
        if (lane == 0) it = atomicAdd(next_item, 1);
        it = __shfl_sync(0xffffffffu, it, 0);
        if (it >= total_items) break;
        int lo = 0, hi = n_probe;   // item_off[lo] <= it < item_off[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (item_off[mid] <= it) lo = mid; else hi = mid;
        }
        const int ar = a0 + lo;
        const uint32_t size = m_size[lo], pos_base = m_pos[lo];
        if (lo != cur) {
            const uint4* tsrc = reinterpret_cast<const uint4*>(a.qtabs) + (static_cast<size_t>(q) * a.ma + ar) * M;
            __syncwarp();   // every lane is done with the previous table
            if (lane < M) wtab[lane] = __ldg(tsrc + lane);
            __syncwarp();
            cur = lo;
        }
        const uint8_t* base = a.codes + m_off[lo] * kSbBytes;
        const uint32_t n_sb = (size + kSbVec - 1) / kSbVec;
        const uint32_t sb0 = static_cast<uint32_t>(it - item_off[lo]) * a.sb_per_item;
        const uint32_t sb1 = min(sb0 + a.sb_per_item, n_sb);
        for (uint32_t sb = sb0; sb < sb1; ++sb) {
            const uint32_t bound = static_cast<uint32_t>(min(*wl.bound, load_shared_bound(sbound) + 1));
            GroupAcc g;
            const uint4* src = reinterpret_cast<const uint4*>(base + static_cast<size_t>(sb) * kSbBytes) + lane;
            uint4 w[kQuads];
#pragma unroll
            for (int qd = 0; qd < kQuads; ++qd) w[qd] = __ldg(src + qd * 32);
#pragma unroll
            for (int p = 0; p < M / 2; ++p) {   // pairs of sub-quantisers, tables read from shared memory (broadcast)
                const uint4& wq = w[p >> 1];
                scan_pair(p == 0, (p & 1) ? wq.z : wq.x, (p & 1) ? wq.w : wq.y, wtab[2 * p], wtab[2 * p + 1], g, pk, bound);
            }
            const bool mine = any_below(g);
            if (__any_sync(0xffffffffu, mine)) {
                const int before = *wl.count;
                __syncwarp();
                if (mine) emit_candidates(g, bound, sb * kSbVec + lane * 8, size, pos_base, ar, wl, hist);
                __syncwarp();
                const int now = *wl.count;
                hist_update(hist, hist_total, hist_next, now - before, a.r, lane, sbound);
                if (now >= compact_at) {
                    // cheap first: drop what the shared bound has overtaken; sort only if that frees little
                    wl.filter(load_shared_bound(sbound), lane);
                    if (*wl.count >= compact_at / 2) wl.compact(a.cap, a.r, lane, sbound);
                }
            }
        }
    }
    wl.filter(load_shared_bound(sbound), lane);
    wl.compact(a.cap, a.r, lane, sbound);
    store_list(wl, a.lists + (static_cast<size_t>(q) * a.n_lists + blockIdx.x * NW + warp) * a.r, a.r, lane);
}

// ------------------------------------------------------------------------------------------
// Shared-bound seed: the r-th smallest int8 distance among the keep-prefixes of the probed
// partitions.  Those prefix vectors are part of the scanned database, so at least r scanned
// vectors have d <= seed and nothing farther can reach the top r.  Two tiny kernels:
// a 128-bin histogram per query (grid = (splits, queries)) and its scan (grid = queries).
// ------------------------------------------------------------------------------------------
struct PrefixBoundArgs {
    const uint8_t* starts;
    const uint64_t* start_off;
    const uint32_t* start_size;
    const int32_t* assign;     // [nq][ma]
    const int8_t* qtabs;       // [nq][ma][M*16]
    int ma, nsplit;
    unsigned int* hist;        // [nq][128], zeroed
    const uint32_t* owned_size = nullptr;   // "owner computes": [K] sizes on this device; probes of lists held elsewhere have no table here
};

template <int M>
__global__ void __launch_bounds__(256) prefix_hist_kernel(const PrefixBoundArgs a) {
    constexpr int CS = M / 2;
    __shared__ unsigned int hist[128];
    __shared__ __align__(16) int8_t tab[8][M * 16];   // one table per warp
    const int split = blockIdx.x, q = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < 128) hist[tid] = 0;
    __syncthreads();
    // inverted lists: warps take probes round-robin and lanes stride over that probe's short prefix;
    // a flat database (one probe) is strided by the whole CTA
    const bool single = a.ma == 1;
    const uint32_t vstep = single ? 256u : 32u, vfirst = single ? tid : lane;
    for (int ar = single ? 0 : warp; ar < a.ma; ar += single ? 1 : 8) {
        const int p = a.assign[static_cast<size_t>(q) * a.ma + ar];
        const uint32_t n = (a.owned_size && a.owned_size[p] == 0) ? 0u : a.start_size[p];
        const uint32_t v0 = static_cast<uint32_t>(static_cast<uint64_t>(n) * split / a.nsplit);
        const uint32_t v1 = static_cast<uint32_t>(static_cast<uint64_t>(n) * (split + 1) / a.nsplit);
        if (v1 == v0) continue;   // warp-uniform
        __syncwarp();
        const uint4* tsrc = reinterpret_cast<const uint4*>(a.qtabs + (static_cast<size_t>(q) * a.ma + ar) * M * 16);
        for (int i = lane; i < M; i += 32) reinterpret_cast<uint4*>(tab[warp])[i] = __ldg(tsrc + i);
        __syncwarp();
        const uint8_t* codes = a.starts + a.start_off[p] * CS;
        // four vectors per thread and turn, loads first (the prefix streams from L2 / HBM: one dependent load per turn was
        // latency-bound).  Distances of 127 and more are not counted: bin 127 is never read, and nearly every prefix
        // vector lands there (qmax is the r-th smallest prefix distance), i.e. a same-address atomic per vector.
        for (uint64_t v = v0 + vfirst; v < v1; v += 4ull * vstep) {
            uint32_t w[4][CS / 4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint64_t vv = v + static_cast<uint64_t>(u) * vstep;
                if (vv < v1) {
                    if constexpr (CS == 8) {
                        const uint2 c = *reinterpret_cast<const uint2*>(codes + vv * CS);
                        w[u][0] = c.x; w[u][1] = c.y;
                    } else {
                        const uint4 c = *reinterpret_cast<const uint4*>(codes + vv * CS);
                        w[u][0] = c.x; w[u][1] = c.y; w[u][2] = c.z; w[u][3] = c.w;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (v + static_cast<uint64_t>(u) * vstep < v1) {
                    int sum = 0;
#pragma unroll
                    for (int j = 0; j < M; ++j) sum += tab[warp][j * 16 + ((w[u][j >> 3] >> (4 * (j & 7))) & 15u)];
                    if (sum < 127) atomicAdd(&hist[sum], 1u);
                }
            }
        }
    }
    __syncthreads();
    if (tid < 127 && hist[tid]) atomicAdd(&a.hist[static_cast<size_t>(q) * 128 + tid], hist[tid]);
}

__global__ void __launch_bounds__(128) prefix_bound_kernel(const unsigned int* __restrict__ hist, int r,
                                                          int* __restrict__ shared_bound) {
    __shared__ unsigned int h[128];
    const int q = blockIdx.x;
    h[threadIdx.x] = hist[static_cast<size_t>(q) * 128 + threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int cum = 0;
        int b = 126;   // fewer than r prefix vectors below 127: everything below 127 may pass
        for (int d = 0; d < 127; ++d) {
            cum += h[d];
            if (cum >= static_cast<unsigned int>(r)) { b = d; break; }
        }
        shared_bound[q] = b;
    }
}

// ------------------------------------------------------------------------------------------
// Stage D: quantised distance of every vector of one partition for one table — runs the
// same lookup core as the scan kernels and writes min(127, sum).
// ------------------------------------------------------------------------------------------
template <int M>
__global__ void __launch_bounds__(256) dump_distances_kernel(const uint8_t* __restrict__ native, uint32_t size,
                                                             const int8_t* __restrict__ qtab,
                                                             int8_t* __restrict__ out, const PipeK pk) {
    constexpr int kQuads = M / 4, kSbBytes = M * 128;
    const uint32_t sb = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    const uint32_t lane = threadIdx.x & 31;
    if (static_cast<uint64_t>(sb) * kSbVec >= size) return;
    GroupAcc g;
    acc_init(g, 0);
    const uint4* src = reinterpret_cast<const uint4*>(native + static_cast<size_t>(sb) * kSbBytes) + lane;
    const uint4* t = reinterpret_cast<const uint4*>(qtab);
#pragma unroll
    for (int qd = 0; qd < kQuads; ++qd) {
        const uint4 tq[4] = {__ldg(t + 4 * qd), __ldg(t + 4 * qd + 1), __ldg(t + 4 * qd + 2), __ldg(t + 4 * qd + 3)};
        lut_quad(__ldg(src + qd * 32), tq, g, pk);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t pos = sb * kSbVec + lane * 8 + k;
        if (pos < size) out[pos] = static_cast<int8_t>(min(127u, lane_sum(g, k, 0)));
    }
}

// ------------------------------------------------------------------------------------------
// Top-r merge: one CTA per query streams L sorted lists of r keys (optionally with ids) and
// keeps the r smallest in a shared-memory buffer (bitonic sort, bound filter).  Used to merge
// per-warp lists of the scan kernels, per-split prefix lists, and the shards of a
// multi-GPU database after the NCCL all-gather.
// Outputs (any may be null): keys, ids, int8 distances, count of real entries.
// Label resolution (IVF): id = labels[label_off[assign[q][rank]] + (pos - pos_base[part])]
// when `labels` is given; otherwise id = in_ids[...] if given, else the key's low 32 bits.
// ------------------------------------------------------------------------------------------
struct MergeArgs {
    const uint64_t* in_keys;   // [nq][L][r]  (or [L][nq][r] when shard_major)
    const uint32_t* in_ids;    // same shape or null
    int L, r, nq, shard_major;
    uint64_t* out_keys;        // [nq][r]
    uint32_t* out_ids;
    int8_t* out_dists;
    int32_t* out_counts;
    float* out_rth_value;      // [nq]: float bits of the r-th key >> 32, FLT_MAX if < r keys
    const int* init_bound;     // [nq] or null: only keys with distance <= init_bound[q] can be in the result (the scan's shared bound)
    // label resolution
    const uint32_t* labels;
    const uint64_t* label_off;       // [K]
    const uint32_t* part_pos_base;   // [K]
    const int32_t* assign;           // [nq][ma]
    int ma;
};

constexpr int kMergeCap = 2048;      // supports r <= 1024
constexpr int kMergeThreads = 256;

__global__ void __launch_bounds__(kMergeThreads) merge_lists_kernel(const MergeArgs a) {
    __shared__ uint64_t keys[kMergeCap];
    __shared__ uint32_t vals[kMergeCap];   // payload: index of the source slot
    __shared__ int count;
    __shared__ unsigned long long bound_key;
    const int q = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < kMergeCap; i += kMergeThreads) { keys[i] = kEmptyKey; vals[i] = 0; }
    if (tid == 0) {
        count = 0;
        bound_key = a.init_bound ? (static_cast<unsigned long long>(a.init_bound[q] + 1) << 48) : kEmptyKey;
    }
    __syncthreads();
    const int total = a.L * a.r;
    const int step = kMergeCap / 2;   // inputs per round; buffer holds <= r <= cap/2 before a round
    int first_base = 0;
    if (a.shard_major && !a.init_bound) {
        // Shard merge: every input list is sorted, so the r-th key of a FULL list bounds the result (that shard alone holds
        // r keys at or below it); with the smallest such key as the filter ~r-and-a-few of the L * r keys survive and the
        // sort shrinks from 1024 to 128-256 slots (config 5 on 8 GPUs: 0.57 ms per 10 000-query batch before).
        for (int l = tid; l < a.L; l += kMergeThreads) {
            const unsigned long long last = a.in_keys[(static_cast<size_t>(l) * a.nq + q) * a.r + a.r - 1];
            if (last != kEmptyKey) atomicMin(&bound_key, last + 1);   // strict filter below: keep keys <= last
        }
        __syncthreads();
    }
    if (a.init_bound && !a.shard_major && total > kMergeCap) {
        // Pre-pass (int8-distance keys, more input than the buffer holds): the scan's shared bound is the r-th distance of
        // ONE CTA's share of the vectors, so nearly every key of every list passes it (148 lists x 100 keys at the 1e9
        // scan: the fast path below overflowed and the streaming rounds cost 73 us per step).  The r-th smallest distance
        // of the union is exact and cheap: a 128-bin histogram of the keys' distances (one shared-memory atomic per
        // distinct value and warp), its scan, and the bound drops to "r keys and the ties of the last distance".
        __shared__ unsigned int dh[129];
        for (int i = tid; i < 129; i += kMergeThreads) dh[i] = 0;
        __syncthreads();
        const unsigned long long bk0 = bound_key;
        constexpr int U = 8;
        for (int base = tid; base - tid < total; base += kMergeThreads * U) {   // warp-uniform trip count (match_any below)
            uint64_t k[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = base + u * kMergeThreads;
                k[u] = i < total ? a.in_keys[(static_cast<size_t>(q) * a.L) * a.r + i] : kEmptyKey;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned d = k[u] < bk0 ? static_cast<unsigned>(k[u] >> 48) & 127u : 128u;
                const unsigned peers = __match_any_sync(0xffffffffu, d);
                if ((tid & 31) == __ffs(peers) - 1) atomicAdd(&dh[d], static_cast<unsigned>(__popc(peers)));
            }
        }
        __syncthreads();
        if (tid == 0) {
            unsigned cum = 0;
            for (int d = 0; d < 128; ++d) {
                cum += dh[d];
                if (cum >= static_cast<unsigned>(a.r)) {
                    const unsigned long long nb = static_cast<unsigned long long>(d + 1) << 48;
                    if (nb < bound_key) bound_key = nb;
                    break;
                }
            }
        }
        __syncthreads();
    }
    if (a.init_bound || a.shard_major) {
        // Fast path: with the scan's final shared bound as the filter only r-and-a-few keys survive, so one pass
        // without per-round barriers collects them all (148 lists x 100 keys: 15 rounds of load -> barrier -> barrier
        // cost 87 us per step of the 1e9 scan); if they do not fit the buffer the streaming rounds below start over.
        const unsigned long long bk0 = bound_key;
        constexpr int U = 8;   // loads in flight per thread: the keys come from L2 / HBM, one dependent load per turn is latency-bound
        for (int base = tid; base < total; base += kMergeThreads * U) {
            uint64_t k[U];
            size_t src[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = base + u * kMergeThreads;
                const int l = i / a.r, e = i % a.r;
                src[u] = a.shard_major ? (static_cast<size_t>(l) * a.nq + q) * a.r + e : (static_cast<size_t>(q) * a.L + l) * a.r + e;
                k[u] = i < total ? a.in_keys[src[u]] : kEmptyKey;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (k[u] < bk0) {
                    const int slot = atomicAdd(&count, 1);
                    if (slot < kMergeCap) { keys[slot] = k[u]; vals[slot] = static_cast<uint32_t>(src[u]); }
                }
            }
        }
        __syncthreads();
        const int c = count;
        __syncthreads();
        if (c <= kMergeCap) {
            int n_sort = 64;
            while (n_sort < c) n_sort <<= 1;
            bitonic_sort_u64_u32(keys, vals, n_sort, tid, kMergeThreads, BlockSync());
            if (tid == 0) count = min(c, a.r);
            __syncthreads();
            first_base = total;   // nothing left to stream
        } else {
            for (int i = tid; i < kMergeCap; i += kMergeThreads) { keys[i] = kEmptyKey; vals[i] = 0; }
            if (tid == 0) count = 0;
            __syncthreads();
        }
    }
    for (int base = first_base; base < total; base += step) {
        const unsigned long long bk = bound_key;
        for (int i = base + tid; i < min(base + step, total); i += kMergeThreads) {
            const int l = i / a.r, e = i % a.r;
            const size_t src = a.shard_major ? (static_cast<size_t>(l) * a.nq + q) * a.r + e
                                             : (static_cast<size_t>(q) * a.L + l) * a.r + e;
            const uint64_t k = a.in_keys[src];
            if (k < bk) {
                const int slot = atomicAdd(&count, 1);
                keys[slot] = k;
                vals[slot] = static_cast<uint32_t>(src);
            }
        }
        __syncthreads();
        const int c = count;
        __syncthreads();   // every thread has read `count` before anyone pushes again (uniform decision)
        if (c > kMergeCap - step || base + step >= total) {
            int n_sort = 64;   // only the occupied power-of-two prefix needs sorting (the rest is empty)
            while (n_sort < c) n_sort <<= 1;
            bitonic_sort_u64_u32(keys, vals, n_sort, tid, kMergeThreads, BlockSync());
            const int n = min(c, a.r);
            for (int i = a.r + tid; i < n_sort; i += kMergeThreads) keys[i] = kEmptyKey;
            if (tid == 0) { count = n; bound_key = (n == a.r) ? keys[a.r - 1] : kEmptyKey; }
            __syncthreads();
        }
    }
    if (total == 0) __syncthreads();
    const int n = count;
    for (int i = tid; i < a.r; i += kMergeThreads) {
        const uint64_t k = keys[i];
        const bool real = i < n;
        const size_t o = static_cast<size_t>(q) * a.r + i;
        if (a.out_keys) a.out_keys[o] = real ? k : kEmptyKey;
        if (a.out_dists) a.out_dists[o] = real ? static_cast<int8_t>(k >> 48) : static_cast<int8_t>(127);
        if (a.out_ids) {
            uint32_t id = 0;
            if (real) {
                const uint32_t pos = static_cast<uint32_t>(k);
                if (a.labels) {
                    const uint32_t rank = static_cast<uint32_t>(k >> 32) & 0xffffu;
                    const int p = a.assign[static_cast<size_t>(q) * a.ma + rank];
                    id = a.labels[a.label_off[p] + (pos - a.part_pos_base[p])];
                } else if (a.in_ids) {
                    id = a.in_ids[vals[i]];
                } else {
                    id = pos;
                }
            }
            a.out_ids[o] = id;
        }
    }
    if (tid == 0) {
        if (a.out_counts) a.out_counts[q] = n;
        if (a.out_rth_value)
            a.out_rth_value[q] = (n == a.r) ? __uint_as_float(static_cast<uint32_t>(keys[a.r - 1] >> 32))
                                            : 3.402823466e+38f;
    }
}

}  // namespace qadc
