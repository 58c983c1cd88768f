/* qadc_b200.h — C ABI of the B200-native Quick ADC search path (libqadc_b200.so).
 *
 * The reference (technicolor-research/quick-adc) has no FFI: its search path sits behind
 * two compile-time concepts, the Scanner (scanner_4, db_query_4.cpp:73-310) and the Engine
 * (nns_engine / nns_engine_batch, query_common.hpp:149-309).  The kept C++ host mirror in
 * quick-adc_b200/host/ implements those concepts on top of the entry points below, which
 * are exactly what a maintainer of the reference would bind.
 *
 * Conventions: plain C types, caller-allocated outputs, every function returns 0 on
 * success or a negative QADC_E* code (message via qadc_last_error); the reference's
 * behaviour on the same conditions is `std::cerr << ...; std::exit(1)`, which the host
 * mirror reproduces.  One context = one GPU = one host thread at a time.  There is no CPU
 * fallback: every entry point fails with QADC_ECUDA when no sm_100 device is usable.
 *
 * Data formats (all [probe]-verified against the reference):
 *   codes      row-major, m/2 bytes per vector, code[b] = idx[2b] | idx[2b+1] << 4
 *              (quantizers.hpp:49-68) — what base_db::get_partition returns
 *   codebooks  m x 16 x (dim/m) floats, flat (quantizers.hpp:160-168)
 *   qtables    int8, [query][probe][m][16], entry c at byte c (db_query_4.cpp:65-68)
 *   tables     float, [query][probe][m][16] (query_common.hpp:235-237)
 *   ids        flat: position in the database; IVF: labels[partition][position]
 *
 * Result rule (the stated deterministic tie-break, SURVEY §8c Stage S): for each query the
 * r smallest records under the total order (distance, probe_rank, position) among scanned
 * vectors with distance < 127, sorted by that order, padded with (id 0, distance 127) —
 * the sentinel the reference seeds its heap with (db_query_4.cpp:276).  Per-vector
 * distances are min(127, sum_j qtable[j][code_j]), bit-exact with scan_avx_4
 * (simd_scan.hpp:125-187).
 */
#ifndef QADC_B200_H
#define QADC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QADC_ABI_VERSION 1

enum {
    QADC_OK = 0,
    QADC_EINVAL = -1,   /* bad argument / unsupported configuration (reference: exit(1)) */
    QADC_ECUDA = -2,    /* CUDA error or no sm_100 device */
    QADC_ESTATE = -3,   /* call order violated (e.g. search before finalize) */
    QADC_EBOUND = -4,   /* "Max quantization bound too high" (db_query_4.cpp:271-274):
                           fewer than r vectors in the probed keep-prefixes */
    QADC_ENOMEM = -5
};

typedef struct qadc_ctx qadc_ctx;

/* Per-batch phase times in microseconds, the reference's query_metrics fields
 * (query_common.hpp:21-56), measured with CUDA events on the context's stream. */
typedef struct qadc_metrics {
    double index_us;   /* coarse assignment + residuals      (query_common.hpp:284-286) */
    double rotate_us;  /* OPQ rotation                       (:288-289) */
    double table_us;   /* float lookup tables                (:291-297) */
    double scan_us;    /* prefix scan + quantise + scan + top-r (:299-303) */
    double h2d_us, d2h_us; /* host<->device copies of qadc_search (not in the reference) */
} qadc_metrics;

int qadc_abi_version(void);

/* ---- lifecycle ------------------------------------------------------------------------ */
/* device: CUDA ordinal.  stream: a cudaStream_t to run on (e.g. torch's current stream),
 * or NULL for a context-owned stream. */
int qadc_create(int device, void* stream, qadc_ctx** out);
void qadc_destroy(qadc_ctx* ctx);
/* Message of the last failing call on ctx (ctx == NULL: last qadc_create failure). */
const char* qadc_last_error(const qadc_ctx* ctx);

/* ---- quantisers ----------------------------------------------------------------------- */
/* Replaces base_pq / opq state (quantizers.hpp:96-168, :248-301).  The Quick ADC database needs bits = 4 and m in
 * {16, 32} (get_simd_scan_func_epi8, db_query_4.cpp:23-35); the plain ADC entry points (qadc_adc_load / qadc_adc_search)
 * and qadc_encode also take (4,8) (8,8) (16,8) (2,16) (4,16) (8,16) (get_scan_func, query_common.hpp:122-147).
 * codebooks: m x 2^bits x dim/m floats.  rotation: dim*dim row-major (opq::rotation) or NULL for a plain PQ.  Host pointers. */
int qadc_set_pq(qadc_ctx* ctx, int dim, int m, int bits, const float* codebooks,
                const float* rotation);
/* Replaces index_db::centroids (databases.hpp:176-189): K*dim floats. K == 0 / never
 * called = flat database (flat_db: one partition, assign = 0, residual = query). */
int qadc_set_coarse(qadc_ctx* ctx, int K, const float* centroids);

/* ---- database upload = scanner_4::prepare_database (db_query_4.cpp:98-228) ------------ */
/* sizes[p] = vectors of partition p held by THIS context (0 allowed: "Partition is empty",
 * db_query_4.cpp:113-116).  has_labels: IVF partitions carry labels, flat ones do not
 * (mixing is an error in the reference too, :118-124). */
int qadc_begin_database(qadc_ctx* ctx, int partition_count, const uint32_t* sizes,
                        int has_labels);
/* Copies vectors [first, first+count) of partition part_i (row-major codes and, if the
 * database has labels, their labels) and re-lays them out on the device.  `first` must be
 * a multiple of 256.  on_device != 0: codes/labels are device pointers.  The caller keeps
 * ownership (the reference's scanner then calls db.free_partition, db_query_4.cpp:190). */
int qadc_upload_codes(qadc_ctx* ctx, int part_i, uint32_t first, uint32_t count,
                      const uint8_t* codes, const uint32_t* labels, int on_device);
/* The whole database in one call: `codes` / `labels` hold the partitions back to back in partition
 * order (sum of sizes vectors), e.g. an inverted index kept as one sorted array.  One copy and one
 * re-layout launch per run of partitions instead of one per partition (65 536 lists: seconds -> ms). */
int qadc_upload_database(qadc_ctx* ctx, const uint8_t* codes, const uint32_t* labels, int on_device);
/* Same from one host pointer per partition, as base_db::get_partition hands them out
 * (databases.hpp:120-134, :233-250): part_codes[p] / part_labels[p] may be NULL where sizes[p] == 0. */
int qadc_upload_partitions(qadc_ctx* ctx, const uint8_t* const* part_codes, const uint32_t* const* part_labels);
/* Sharded databases only: position of this context's first vector inside the full
 * partition (flat database split across GPUs), default 0. */
int qadc_set_position_base(qadc_ctx* ctx, int part_i, uint32_t pos_base);
/* Sharded databases only: explicit keep-prefix of partition part_i (row-major codes of the
 * first `count` vectors of the FULL partition), replicated on every shard so that all
 * shards derive identical quantisation bounds.  Overrides the prefix finalize derives. */
int qadc_set_prefix(qadc_ctx* ctx, int part_i, const uint8_t* codes, uint32_t count,
                    int on_device);
/* Same for every partition at once: `codes` holds the prefixes back to back in partition order,
 * counts[p] vectors each (0 = derive that partition's prefix from the uploaded codes). */
int qadc_set_prefixes(qadc_ctx* ctx, const uint8_t* codes, const uint32_t* counts, int on_device);
/* keep: fraction of each partition scanned in float to bound the quantiser,
 * starts_sizes[p] = max(1, (unsigned)(size * keep)) in float32 (db_query_4.cpp:125-126). */
int qadc_finalize(qadc_ctx* ctx, float keep);

/* ---- search: the call engine_gpu makes once per query batch --------------------------- */
/* Host buffers in, host buffers out (H2D/D2H inside).  queries: nq*dim.  ma: probes per
 * query (must be 1 for a flat database).  out_ids/out_dists: nq*r; out_counts[q]: number
 * of real (non-sentinel) entries.  metrics may be NULL.
 * Replaces nns_engine_batch::batch_process_queries + scanner_4::query_scan per query
 * (query_common.hpp:194-242, db_query_4.cpp:245-309). */
int qadc_search(qadc_ctx* ctx, const float* queries, int nq, int ma, int r, uint32_t* out_ids,
                int8_t* out_dists, int32_t* out_counts, qadc_metrics* metrics);
/* Same with device-resident buffers, asynchronous on the context's stream.  out_keys
 * (nq*r uint64, optional) receives the canonical sort keys
 * (distance << 48 | probe_rank << 32 | position) used to merge shards. */
int qadc_search_device(qadc_ctx* ctx, const float* d_queries, int nq, int ma, int r,
                       uint32_t* d_ids, int8_t* d_dists, int32_t* d_counts, uint64_t* d_keys);
/* Same with the coarse assignment supplied by the caller (d_assign: nq*ma cell indices in probe
 * order, on the device) instead of computed: the second half of a sharded search, after the
 * ranks have exchanged their partial assignments (below). */
int qadc_search_assigned_device(qadc_ctx* ctx, const float* d_queries, const int32_t* d_assign, int nq,
                                int ma, int r, uint32_t* d_ids, int8_t* d_dists, int32_t* d_counts,
                                uint64_t* d_keys);
/* "Owner computes": the per-query table pipeline of SHARDED inverted lists, split around one exchange, so that no
 * shard builds tables or scans keep-prefixes for lists it does not hold (scanner_4::query_scan, db_query_4.cpp:245-284,
 * whose bounds are one qmin/qmax per QUERY over all ma probes).  A shard answers for the partitions it owns
 * (qadc_set_owned_partitions); it needs no replica of the other lists' prefixes.
 *   1. qadc_tables_local_device: float tables of the OWNED probes of every query (kept in the context's scratch) and
 *      this shard's share of the bounds, d_local[q][0] = min entry of those tables (FLT_MAX if it owns none),
 *      d_local[q][1..r] = its r smallest keep-prefix distances (scan_4, query_common.hpp:59-90; FLT_MAX-padded).
 *      d_local: nq * (r + 1) floats on the device.  nq <= 32768 per call.
 *   2. the caller all-gathers d_local over the shards: d_gathered[g][q][r + 1], in any shard order.
 *   3. qadc_search_bounded_device: qmin = min over shards, qmax = r-th smallest of the union (both order-independent
 *      selections of values computed with the unsharded arithmetic, hence bit-identical to the unsharded bounds), int8
 *      tables of the owned probes (QuantizerMAX, db_query_4.cpp:37-71), scan of the owned lists, per-shard canonical top-r
 *      (same outputs as qadc_search_device; merge the shards with qadc_merge_shards_device).  Must follow step 1 of the
 *      same batch on the same context.  A query whose union holds fewer than r prefix vectors reports QADC_EBOUND at the
 *      next qadc_synchronize, like the unsharded search. */
/* Which partitions this shard answers for (uint8 mask [partition_count], after qadc_begin_database; default: the
 * partitions given a non-zero size).  Every partition must be owned by exactly one shard — including partitions that are
 * empty everywhere, whose tables still take part in the query's qmin (db_query_4.cpp:256-260 runs over all ma tables). */
int qadc_set_owned_partitions(qadc_ctx* ctx, const uint8_t* owned);
int qadc_tables_local_device(qadc_ctx* ctx, const float* d_queries, const int32_t* d_assign, int nq, int ma, int r,
                             float* d_local);
int qadc_search_bounded_device(qadc_ctx* ctx, const float* d_gathered, int G, int nq, int ma, int r, uint32_t* d_ids,
                               int8_t* d_dists, int32_t* d_counts, uint64_t* d_keys);
/* Sharded coarse assignment (index_db::assign_compute_residuals_mutiple -> find_k_neighbors,
 * databases.hpp:213-231, neighbors.cpp:30-76, split over the GPUs of one box): every rank ranks
 * the queries against the cells [c_first, c_first + c_count) only and returns its ma best as keys
 * (float distance bits << 32 | global cell index), ascending, padded with ~0 when c_count < ma
 * (d_out_keys: nq*ma).  After an all-gather, qadc_coarse_merge_device keeps the ma smallest of the
 * G*ma keys of every query (d_keys laid out [G][nq][ma]) and writes their cell indices to d_assign
 * (nq*ma) — the same assignment, bit for bit, as the unsharded search computes. */
int qadc_coarse_partial_device(qadc_ctx* ctx, const float* d_queries, int nq, int ma, int c_first,
                               int c_count, uint64_t* d_out_keys);
int qadc_coarse_merge_device(qadc_ctx* ctx, const uint64_t* d_keys, int G, int nq, int ma, int32_t* d_assign);
/* Waits for the context's stream and reports deferred device-side conditions of earlier
 * asynchronous calls (QADC_EBOUND from qadc_search_device). */
int qadc_synchronize(qadc_ctx* ctx);
/* Number of kernels launched by the last search call (bench accounting). */
int qadc_last_launch_count(const qadc_ctx* ctx);
/* Device time in milliseconds of the scan kernel(s) of the last search call (CUDA events on
 * the context's stream around the scan launches; synchronises the stream). */
int qadc_last_scan_ms(qadc_ctx* ctx, float* ms);
/* Same for the last n search calls (oldest first, at most 64): no host synchronisation is
 * needed inside a timed region.  Returns how many values were written, or a negative code. */
int qadc_scan_ms_history(qadc_ctx* ctx, float* ms, int n);

/* Merge the local top-r lists of G shards (device buffers laid out [G][nq][r], as an NCCL
 * all-gather leaves them) into the global top-r under the same total order. */
int qadc_merge_shards_device(qadc_ctx* ctx, const uint64_t* d_keys, const uint32_t* d_ids,
                             int G, int nq, int r, uint32_t* d_out_ids, int8_t* d_out_dists,
                             int32_t* d_out_counts, uint64_t* d_out_keys);

/* ---- one database over several GPUs of one box, one host process ----------------- */
/* The reference scans with one thread on one CPU (query_common.hpp:351-365); this is what replaces it when
 * the database is sharded: a qadc_multi owns one context per listed device, shards the database itself
 * (flat: contiguous runs of 256-vector blocks; inverted lists: whole lists, longest first; the keep-prefixes
 * replicated on every shard), and per query batch exchanges the per-shard coarse candidates (inverted lists)
 * and the per-shard top-r lists with NCCL all-gathers over NVLink before merging them on the first device.
 * Results are bit-identical to the single-GPU ones.  A device ordinal listed more than once = several virtual
 * shards on one GPU (NCCL rejects that; the gather is then done with same-device copies) — for tests. */
typedef struct qadc_multi qadc_multi;
int qadc_multi_create(const int* devices, int n, qadc_multi** out);
void qadc_multi_destroy(qadc_multi* m);
const char* qadc_multi_last_error(const qadc_multi* m);   /* m == NULL: last qadc_multi_create failure */
int qadc_multi_device_count(const qadc_multi* m);
int qadc_multi_uses_nccl(const qadc_multi* m);
qadc_ctx* qadc_multi_context(qadc_multi* m, int g);        /* shard g's context (options, parity entry points) */
int qadc_multi_set_pq(qadc_multi* m, int dim, int mm, int bits, const float* codebooks, const float* rotation);
int qadc_multi_set_coarse(qadc_multi* m, int K, const float* centroids);
/* scanner_4::prepare_database (db_query_4.cpp:98-228) for all shards: sizes[p], part_codes[p] (row-major) and
 * part_labels[p] (inverted lists; NULL for a flat database) as base_db::get_partition hands them out. */
int qadc_multi_load(qadc_multi* m, int partition_count, const uint32_t* sizes, const uint8_t* const* part_codes,
                    const uint32_t* const* part_labels, float keep);
/* = qadc_search over the sharded database (host buffers in and out). */
int qadc_multi_search(qadc_multi* m, const float* queries, int nq, int ma, int r, uint32_t* out_ids, int8_t* out_dists,
                      int32_t* out_counts, qadc_metrics* metrics);

/* ---- parity entry points (host buffers) ------------------------------------------------ */
/* Stage T: assignment, float tables, bounds and int8 tables for a batch.  assign_in
 * (nq*ma, optional) injects a coarse assignment instead of computing it.  Any output may
 * be NULL.  Returns QADC_EBOUND if some query's prefix holds fewer than r vectors. */
int qadc_build_tables(qadc_ctx* ctx, const float* queries, int nq, int ma, int r,
                      const int32_t* assign_in, int32_t* out_assign, float* out_tables,
                      float* out_qmin, float* out_qmax, int8_t* out_qtables);
/* Stage S with injected int8 tables (the bit-exact stage): assign nq*ma, qtables
 * nq*ma*m*16 with every entry in [0,127]. */
int qadc_scan_with_tables(qadc_ctx* ctx, const int32_t* assign, const int8_t* qtables, int nq,
                          int ma, int r, uint32_t* out_ids, int8_t* out_dists,
                          int32_t* out_counts);
/* Stage D: the quantised distance of every vector of partition part_i for one int8 table
 * (m*16), out[size]. */
int qadc_dump_distances(qadc_ctx* ctx, int part_i, const int8_t* qtable, int8_t* out);
/* The device layout read back as row-major codes (layout round-trip test). */
int qadc_download_codes(qadc_ctx* ctx, int part_i, uint8_t* out_codes);

/* ---- "next" row N1: PQ encoder -------------------------------------------------------------- */
/* Replaces base_pq::encode_multiple_vectors + multiple_set_bits_4 (quantizers.hpp:49-68,
 * :222-245) and, with a coarse quantiser set, index_db::assign_single_compute_residuals
 * (databases.hpp:252-268): rotates (OPQ; with inverted lists the residual is what is rotated) and encodes `count` host vectors
 * (count*dim floats) into row-major codes (count*m*bits/8 bytes; 8- / 16-bit quantisers: one byte / one little-endian uint16
 * per sub-quantiser, multiple_set_bits_native<T>, quantizers.hpp:36-47).  out_assign (count, may be
 * NULL) receives the coarse cell of every vector when the context has a coarse quantiser; the
 * code is then that of the residual.  Nearest centroid by direct squared distance, first minimum
 * on ties. */
int qadc_encode(qadc_ctx* ctx, const float* vectors, uint32_t count, int32_t* out_assign,
                uint8_t* out_codes);

/* ---- plain ADC: the reference's db_query tool ("next" row: scanner_simple, db_query.cpp:17-46;
 * scan_standard<uint8_t,NSQ> / scan_4<NSQ>, query_common.hpp:59-147) ---------------------------- */
/* Row-major codes of the whole database, partitions concatenated: partition p holds the vectors
 * [offsets[p], offsets[p+1]) (partition_count + 1 offsets; 1 partition for a flat database, K for
 * inverted lists, which also need `labels`).  Code size m*bits/8 bytes; 4-bit codes two per byte,
 * low nibble first (quantizers.hpp:49-68).  Works for every quantiser qadc_set_pq accepts:
 * (16,4) (32,4) (4,8) (8,8) (16,8) (2,16) (4,16) (8,16) — every pair of get_scan_func (query_common.hpp:122-147);
 * 16-bit codes are one little-endian uint16 per sub-quantiser.
 * Independent of the Quick ADC database (qadc_begin_database .. qadc_finalize). */
int qadc_adc_load(qadc_ctx* ctx, int partition_count, const uint64_t* offsets, const uint8_t* codes,
                  const uint32_t* labels);
/* Float ADC search, host buffers: out_ids / out_dists nq*r, ascending by (distance, probe rank,
 * position) — the entries the reference's heap ends with (strict `candidate < max` test in scan
 * order); out_counts[q] real entries, the rest are (0, FLT_MAX) like the reference's pre-filled
 * heap slots (db_query.cpp:27-30).  Distances are summed in sub-quantiser order in float32. */
int qadc_adc_search(qadc_ctx* ctx, const float* queries, int nq, int ma, int r, uint32_t* out_ids,
                    float* out_dists, int32_t* out_counts);

/* ---- tuning knobs (bench / tests) ------------------------------------------------------ */
/* key: "flat_qb" (queries per pass of the flat scan: 1,2,4,8), "flat_chunks" (CTAs along
 * the database, 0 = auto), "flat_filter" (1 = default: clamped byte-lane pre-filter in front of the exact
 * lookup core of the one-query-per-pass flat scan, 0 = exact core only; same results), "ivf_fused" (1 = default:
 * one kernel per query batch builds the float tables of an inverted-list search in shared memory, bounds and
 * quantises them there; 0 = the separate table / prefix / quantise kernels; same results), "ivf_sb_per_item" (256-vector blocks per work item of the inverted-list
 * scan, default 8), "flat_ring" (1 = default: per-warp TMA rings; 0 = the CTA-wide ring of round 1), "flat_share"
 * (1 = default: the CTAs that scan different chunks of a flat database for the same query pool their candidates in a
 * global histogram, so that the query's bound is that of everything scanned so far; 0 = every CTA derives it from its
 * own chunk; same results), "flat_prep" (1 = default: flat databases whose keep-prefix has at least 131 072 vectors
 * select the candidates of the float prefix scan with int8 lower bounds; 0 = the plain float scan of the whole prefix;
 * same qmax, tables and results), "flat_seed" (1 = default: the scan's bound starts at a value derived from the
 * keep-prefix; 0 = at 126; same results), "time_scan" (1: record CUDA events around the scan kernel for
 * qadc_last_scan_ms).  Every setting returns the same results; the knobs exist for A/B timing and for the parity
 * tests that compare the paths.  Unknown key -> QADC_EINVAL. */
int qadc_set_option(qadc_ctx* ctx, const char* key, long value);

#ifdef __cplusplus
}
#endif
#endif /* QADC_B200_H */
