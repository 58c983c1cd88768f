// qadc_device.cuh — device-side building blocks of the B200 Quick ADC scan (sm_100a).
//
// Device code layout ("nibble planes"), the GPU counterpart of the reference's transposed
// 16-vector blocks (simd_layout.hpp:41-65):
//   superblock = 256 consecutive vectors = M/4 quads x 32 lanes x 16 bytes (M*128 bytes)
//   the uint4 at (quad q, lane l) holds 4 words, word s = sub-quantiser j = 4q+s,
//   nibble k of the word = centroid index of vector 8*l + k for that sub-quantiser.
// A warp reads one quad of a superblock with one conflict-free 128-bit access per lane,
// and the low/high 16 bits of a word are directly the PRMT selectors of 4 vectors.
//
// Lookup (the GPU analogue of vpshufb in scan_avx_4, simd_scan.hpp:157-173): the 16-entry
// int8 table of a sub-quantiser is one uint4 {T0,T1,T2,T3}; every entry is in [0,127]
// (QuantizerMAX, db_query_4.cpp:44-55), so PRMT's sign-replicate mode (selector bit 3)
// returns 0 for them:
//   lo = prmt(T0, T1, w)               -> T[idx] if idx < 8 else 0
//   hi = prmt(T2, T3, w ^ 0x88888888)  -> T[idx] if idx >= 8 else 0
// Two sub-quantisers are added in byte lanes (<= 254, no carry), split into even/odd bytes
// and accumulated in 16-bit lanes that start at 0x8000 - bound, so the top bit of a lane is
// set iff sum >= bound.  Signed saturation of the reference (vpaddsb on values in [0,127]) equals
// min(127, sum), and only sums < bound <= 127 are ever selected.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace qadc {

constexpr int kSbVec = 256;              // vectors per superblock
constexpr uint64_t kEmptyKey = ~0ull;    // sorts after every real key

__host__ __device__ constexpr int sb_bytes(int m) { return m * 128; }

// canonical sort key (SURVEY §8c Stage S): distance, probe rank, position
__host__ __device__ __forceinline__ uint64_t make_key(uint32_t d, uint32_t probe_rank, uint32_t pos) {
    return (static_cast<uint64_t>(d) << 48) | (static_cast<uint64_t>(probe_rank) << 32) | pos;
}

// ---- PTX helpers: mbarrier + TMA bulk copy ---------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// 32-bit shared-address variants (no generic-address arithmetic in the hot loop)
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
// One lane of a converged warp (elect.sync): unlike `lane == 0` the compiler knows the branch holds a single thread,
// so a TMA issue inside it needs no ELECT / R2UR.BROADCAST loop around the uniform-register operands.
__device__ __forceinline__ bool elect_one() {
    uint32_t leader;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(leader));
    return leader != 0;
}
// Orders the TMA refill of a ring slot behind the shared-memory loads that read the slot.  Returns 0, computed FROM the
// last loaded word (times a run-time zero the compiler cannot see through), to be added to the refill's byte count: the
// refill's operands then depend on the load, so the instruction stream cannot reach the refill while a load still sits in
// the LSU queue.  (A storm of 128-bit table loads of the other warps can hold it there longer than an L2-hit TMA copy
// takes: the batched kernel then saw code words of the NEXT superblock — found by the variant-agreement test.)
__device__ __forceinline__ uint32_t zero_after_loads(uint32_t last_loaded_word, uint32_t runtime_zero) {
    uint32_t d;
    asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(d) : "r"(last_loaded_word), "r"(runtime_zero));
    return d;
}
// keeps a loop-invariant value in a register instead of letting the compiler rematerialise it
__device__ __forceinline__ void pin(uint32_t& v) { asm volatile("" : "+r"(v)); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- the lookup-accumulate core ----------------------------------------------------------
// Raw prmt.b32: bit 3 of a selector nibble replicates the SIGN of the selected byte
// (__byte_perm masks the selector with 0x7777, so it cannot be used here).
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s));
    return d;
}

// Accumulators of one group of 8 vectors: one 32-bit accumulator per vector, fed with IDP.4A (dot
// product with a one-hot byte selector), so neither the byte extraction nor the final test needs
// ALU-pipe masks.  `one`, `neg1` and the selectors are kernel arguments: the compiler cannot fold
// them, so a*one+b is emitted as IMAD on the FMA pipe instead of competing with PRMT/LOP3/SHF for
// the ALU pipe, which is what bounds the scan.  (An earlier variant with packed 16-bit lanes and
// mask/shift extraction was 10 % slower.)
struct GroupAcc {
    uint32_t v[8];   // sum - bound (negative <=> candidate)
};
struct PipeK {
    uint32_t one, neg1, s0, s1, s2, s3;   // 1, -1, and the byte selectors 1, 1<<8, 1<<16, 1<<24
};
// selector of vectors 4..7 = upper half of the word (SHF on the ALU pipe; multiply-high and
// IMAD.WIDE forms on the FMA pipe were measured 10 % / 9 % slower)
__device__ __forceinline__ uint32_t hi16(uint32_t w, const PipeK&) { return w >> 16; }
__device__ __forceinline__ uint32_t fadd(uint32_t a, uint32_t b, const PipeK& k) { return a * k.one + b; }
__device__ __forceinline__ void acc_init(GroupAcc& g, uint32_t bound) {
#pragma unroll
    for (int i = 0; i < 8; ++i) g.v[i] = 0u - bound;
}
// FIRST: the accumulators are not read but started at `start` (= -bound), which saves the
// eight register initialisations per group.
template <bool FIRST>
__device__ __forceinline__ void lut_pair(uint32_t w0, uint32_t w1, const uint4& t0, const uint4& t1, GroupAcc& g,
                                         const PipeK& k, uint32_t start) {
    const uint32_t x0 = w0 ^ 0x88888888u, x1 = w1 ^ 0x88888888u;
    const uint32_t pa = fadd(fadd(prmt(t0.x, t0.y, w0), prmt(t0.z, t0.w, x0), k),
                             fadd(prmt(t1.x, t1.y, w1), prmt(t1.z, t1.w, x1), k), k);
    const uint32_t pb = fadd(fadd(prmt(t0.x, t0.y, hi16(w0, k)), prmt(t0.z, t0.w, hi16(x0, k)), k),
                             fadd(prmt(t1.x, t1.y, hi16(w1, k)), prmt(t1.z, t1.w, hi16(x1, k)), k), k);
    g.v[0] = __dp4a(pa, k.s0, FIRST ? start : g.v[0]); g.v[1] = __dp4a(pa, k.s1, FIRST ? start : g.v[1]);
    g.v[2] = __dp4a(pa, k.s2, FIRST ? start : g.v[2]); g.v[3] = __dp4a(pa, k.s3, FIRST ? start : g.v[3]);
    g.v[4] = __dp4a(pb, k.s0, FIRST ? start : g.v[4]); g.v[5] = __dp4a(pb, k.s1, FIRST ? start : g.v[5]);
    g.v[6] = __dp4a(pb, k.s2, FIRST ? start : g.v[6]); g.v[7] = __dp4a(pb, k.s3, FIRST ? start : g.v[7]);
}
// The same pair with the four selectors of each code word prepared by the caller (Sel): several queries that share
// a pass over the codes share them too (3 ALU-pipe operations per word that are then paid once, not once per query).
struct Sel {
    uint32_t w, x, wh, xh;   // word, word ^ 0x88888888, and their upper halves
};
__device__ __forceinline__ Sel make_sel(uint32_t w, const PipeK& k) {
    const uint32_t x = w ^ 0x88888888u;
    return Sel{w, x, hi16(w, k), hi16(x, k)};
}
template <bool FIRST>
__device__ __forceinline__ void lut_pair_sel(const Sel& s0, const Sel& s1, const uint4& t0, const uint4& t1, GroupAcc& g,
                                             const PipeK& k, uint32_t start) {
    const uint32_t pa = fadd(fadd(prmt(t0.x, t0.y, s0.w), prmt(t0.z, t0.w, s0.x), k),
                             fadd(prmt(t1.x, t1.y, s1.w), prmt(t1.z, t1.w, s1.x), k), k);
    const uint32_t pb = fadd(fadd(prmt(t0.x, t0.y, s0.wh), prmt(t0.z, t0.w, s0.xh), k),
                             fadd(prmt(t1.x, t1.y, s1.wh), prmt(t1.z, t1.w, s1.xh), k), k);
    g.v[0] = __dp4a(pa, k.s0, FIRST ? start : g.v[0]); g.v[1] = __dp4a(pa, k.s1, FIRST ? start : g.v[1]);
    g.v[2] = __dp4a(pa, k.s2, FIRST ? start : g.v[2]); g.v[3] = __dp4a(pa, k.s3, FIRST ? start : g.v[3]);
    g.v[4] = __dp4a(pb, k.s0, FIRST ? start : g.v[4]); g.v[5] = __dp4a(pb, k.s1, FIRST ? start : g.v[5]);
    g.v[6] = __dp4a(pb, k.s2, FIRST ? start : g.v[6]); g.v[7] = __dp4a(pb, k.s3, FIRST ? start : g.v[7]);
}
// One quad = 4 sub-quantisers. FIRST = first quad of a vector (see lut_pair).
template <bool FIRST>
__device__ __forceinline__ void lut_quad_t(const uint4& w, const uint4 (&t)[4], GroupAcc& g, const PipeK& k, uint32_t start) {
    lut_pair<FIRST>(w.x, w.y, t[0], t[1], g, k, start);
    lut_pair<false>(w.z, w.w, t[2], t[3], g, k, start);
}
__device__ __forceinline__ void lut_quad(const uint4& w, const uint4 (&t)[4], GroupAcc& g, const PipeK& k) {
    lut_quad_t<false>(w, t, g, k, 0u);
}
__device__ __forceinline__ bool any_below(const GroupAcc& g) {
    return static_cast<int>(g.v[0] | g.v[1] | g.v[2] | g.v[3] | g.v[4] | g.v[5] | g.v[6] | g.v[7]) < 0;
}
// 16-bit "raw lane" compatible with the packed variant: bit 15 set <=> sum >= bound
__device__ __forceinline__ uint32_t lane_raw(const GroupAcc& g, int k) { return (g.v[k] + 0x8000u) & 0xffffu; }
__device__ __forceinline__ uint32_t lane_sum(const GroupAcc& g, int k, uint32_t bound) { return g.v[k] + bound; }

// ---- pre-filter: clamped tables summed in byte lanes ------------------------------------------
// A vector is a candidate iff S = sum_j T[j][c_j] <= t (t = bound - 1 <= 126).  With F[j][c] =
// min(T[j][c], cap) the sum F_S = sum_j F[j][c_j] <= S, so F_S > t proves S > t: a conservative
// test that needs no widening at all when the byte lanes cannot wrap.  A lane starts at 127 - t and
// receives M entries <= cap, cap = (128 + t) / M, so it ends at 127 - t + F_S <= 255 and its bit 7
// says F_S > t.  Per sub-quantiser word (8 vectors): 1 XOR + 2 SHF + 4 PRMT on the ALU pipe and
// 4 IMAD on the FMA pipe — the 8 IDP.4A + 2 IMAD per word pair of the exact core are gone.  Only
// superblocks in which some vector passes (a few percent once the bound has tightened: the clamp
// loses little because entries above cap already take a large share of t) run the exact core.
struct FiltAcc {
    uint32_t a, b;   // byte lane k of a: vector k (0..3) of the group, of b: vector 4 + k
};
__device__ __forceinline__ int filt_cap(int t, int m) { return (128 + t) / m; }
__device__ __forceinline__ uint32_t filt_start(int t) { return static_cast<uint32_t>(127 - t) * 0x01010101u; }
// a * one + b as an opaque IMAD: written in C++ the compiler factors the multiplications out of the whole
// sum of a superblock and adds the PRMT results with IADD3 on the ALU pipe, the pipe that binds
__device__ __forceinline__ uint32_t madd(uint32_t a, uint32_t b, const PipeK& k) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(k.one), "r"(b));
    return d;
}
#ifndef QADC_FILT_ADD
#define QADC_FILT_ADD 2   // measured on the 1e9 scan: 2 (one accumulator per pipe) 23.6 ms, 1 (IADD3 only) 23.8, 0 (IMAD only) 24.3
#endif
#ifndef QADC_FILT_HI16
#define QADC_FILT_HI16 0   // 0: SHF (ALU pipe), 1: multiply-high by 65536 (FMA pipe), 2: 32x32->64 multiply, upper word
#endif
__device__ __forceinline__ uint32_t filt_hi16(uint32_t w, const PipeK& k) {
#if QADC_FILT_HI16 == 1
    uint32_t d;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(d) : "r"(w), "r"(k.s2));   // s2 = 1 << 16, opaque
    return d;
#elif QADC_FILT_HI16 == 2
    uint64_t d;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(d) : "r"(w), "r"(k.s2));
    return static_cast<uint32_t>(d >> 32);
#else
    return hi16(w, k);
#endif
}
__device__ __forceinline__ void filt_word(uint32_t w, const uint4& f, FiltAcc& g, const PipeK& k) {
    const uint32_t x = w ^ 0x88888888u;
#if QADC_FILT_ADD == 1      // three-input adds on the ALU pipe (half the add instructions, on the busier pipe)
    g.a = g.a + prmt(f.x, f.y, w) + prmt(f.z, f.w, x);
    g.b = g.b + prmt(f.x, f.y, filt_hi16(w, k)) + prmt(f.z, f.w, filt_hi16(x, k));
#elif QADC_FILT_ADD == 2    // one accumulator on each pipe
    g.a = g.a + prmt(f.x, f.y, w) + prmt(f.z, f.w, x);
    g.b = madd(madd(prmt(f.x, f.y, filt_hi16(w, k)), prmt(f.z, f.w, filt_hi16(x, k)), k), g.b, k);
#else
    g.a = madd(madd(prmt(f.x, f.y, w), prmt(f.z, f.w, x), k), g.a, k);
    g.b = madd(madd(prmt(f.x, f.y, filt_hi16(w, k)), prmt(f.z, f.w, filt_hi16(x, k)), k), g.b, k);
#endif
}
// some vector of the group may be a candidate (bit 7 of its lane still clear)
__device__ __forceinline__ bool filt_any(const FiltAcc& g) { return (~(g.a & g.b) & 0x80808080u) != 0u; }

// ---- bounded candidate lists: bitonic sort of u64 keys in shared memory ------------------
// Sorts n (power of two) keys ascending with the threads [0, nthreads) of a group that
// synchronises through `sync()` (a __syncwarp or a named barrier).
template <typename Sync>
__device__ __forceinline__ void bitonic_sort_u64(uint64_t* keys, int n, int tid, int nthreads, Sync sync) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < (n >> 1); i += nthreads) {
                // i-th compare-exchange of this stage: indices (a, a^j) with bit j of a clear
                const int a = ((i & ~(j - 1)) << 1) | (i & (j - 1));
                const int b = a | j;
                const uint64_t ka = keys[a], kb = keys[b];
                const bool up = (a & k) == 0;
                if ((ka > kb) == up) { keys[a] = kb; keys[b] = ka; }
            }
            sync();
        }
    }
}

// Same with a 32-bit payload carried along (shard merge: ids travel with their keys).
template <typename Sync>
__device__ __forceinline__ void bitonic_sort_u64_u32(uint64_t* keys, uint32_t* vals, int n, int tid,
                                                     int nthreads, Sync sync) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < (n >> 1); i += nthreads) {
                const int a = ((i & ~(j - 1)) << 1) | (i & (j - 1));
                const int b = a | j;
                const uint64_t ka = keys[a], kb = keys[b];
                const bool up = (a & k) == 0;
                if ((ka > kb) == up) {
                    keys[a] = kb; keys[b] = ka;
                    const uint32_t va = vals[a]; vals[a] = vals[b]; vals[b] = va;
                }
            }
            sync();
        }
    }
}

struct WarpSync {
    __device__ __forceinline__ void operator()() const { __syncwarp(); }
};
struct NamedSync {   // a subset of the CTA's warps (bar.sync id, nthreads)
    int id, nthreads;
    __device__ __forceinline__ void operator()() const { named_bar_sync(id, nthreads); }
};
struct BlockSync {
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};

// Per-warp candidate list in shared memory.  Holds the r smallest keys seen so far plus
// the candidates appended since the last compaction.  Because a warp visits its vectors in
// increasing canonical order (probe rank, then position), a later vector whose distance
// EQUALS the r-th held distance can never enter the top r, so the pass test is the strict
// `d < bound` of the reference heap (binheap.hpp:93) and ties resolve to the earlier
// position — the stated tie-break.
struct WarpList {
    uint64_t* keys;  // cap entries, unused slots hold kEmptyKey
    int* count;      // entries in use
    int* bound;      // strict local bound: pass iff d < *bound

    __device__ __forceinline__ void push(uint64_t key) {
        const int i = atomicAdd(count, 1);
        keys[i] = key;
    }
    // Drop every entry farther than vmax (the query's shared bound: at least r scanned vectors are
    // at or below it, so nothing farther can reach the top r).  Order-preserving, no sort: one pass
    // with a ballot prefix.  All 32 lanes call it; the list is left unsorted.
    __device__ __forceinline__ void filter(int vmax, int lane) {
        __syncwarp();
        const int cnt = *count;
        int w = 0;
        for (int base = 0; base < cnt; base += 32) {
            const int i = base + lane;
            const uint64_t key = (i < cnt) ? keys[i] : kEmptyKey;
            const bool keep = i < cnt && static_cast<int>(key >> 48) <= vmax;
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            __syncwarp();   // the whole chunk is in registers before anything is written over it
            if (keep) keys[w + __popc(m & ((1u << lane) - 1u))] = key;
            w += __popc(m);
        }
        __syncwarp();
        for (int i = w + lane; i < cnt; i += 32) keys[i] = kEmptyKey;
        if (lane == 0) *count = w;
        __syncwarp();
    }
    // Keep the r smallest (sorted). All 32 lanes call it. Only the occupied power-of-two prefix is
    // sorted.  When the list is full its r-th distance is also published to the query's shared
    // bound (any vector farther than it can no longer be in the top r of the whole scan).
    __device__ __forceinline__ void compact(int cap, int r, int lane, int* shared_bound) {
        __syncwarp();
        const int cnt = *count;
        int n_sort = 32;
        while (n_sort < cnt) n_sort <<= 1;
        n_sort = min(n_sort, cap);
        bitonic_sort_u64(keys, n_sort, lane, 32, WarpSync());
        const int n = min(cnt, r);
        __syncwarp();
        for (int i = r + lane; i < n_sort; i += 32) keys[i] = kEmptyKey;
        if (lane == 0) {
            *count = n;
            if (n == r) {
                const int d = static_cast<int>(keys[r - 1] >> 48);
                *bound = d;
                if (shared_bound) atomicMin(shared_bound, d);
            }
        }
        __syncwarp();
    }
};

}  // namespace qadc
