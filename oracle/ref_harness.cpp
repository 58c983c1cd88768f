// oracle/ref_harness.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Wraps the UNMODIFIED reference sources (compiled where they lie, under
// $QADC_REFERENCE_DIR, default /root/reference) behind a small C ABI so that
//   * tests can pin the C restatement in oracle/qadc_oracle.c against the real code,
//   * tests/golden/ fixtures can be generated from the real code,
//   * bench.py can time the reference's own CPU scan (`--impl reference`,
//     `cpu_baseline.kind == "reference"`).
// The reference has no library boundary: scanner_4, QuantizerMAX, scan_avx_4 live in
// db_query_4.cpp, so that TU is included here with its main() renamed (simd_layout.hpp,
// simd_scan.hpp and query_common.hpp define non-inline functions, so exactly one TU may
// include them).  Nothing in this file restates reference logic except the per-query loop
// body of nns_engine::process_query (query_common.hpp:278-307), which is repeated with
// thread-local buffers because the reference's engines keep one shared scratch buffer and
// its query loop is single-threaded (query_common.hpp:351-365); north_star asks for a
// multi-threaded CPU baseline, so the harness adds `#pragma omp parallel for` over queries.
#define main qadc_ref_main
#include "db_query_4.cpp"
#undef main

#include <cstring>
#include <vector>

#ifdef QADC_REF_NAIVE_BLAS
// Only the (RowMajor, NoTrans, Trans) sgemm form is used by the reference
// (distances.hpp:181,213; quantizers.hpp:296).
extern "C" void qadc_naive_sgemm(enum CBLAS_ORDER, enum CBLAS_TRANSPOSE, enum CBLAS_TRANSPOSE,
                                 int m, int n, int k, float alpha, const float* a, int lda,
                                 const float* b, int ldb, float beta, float* c, int ldc) {
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) {
            float s = 0.f;
            for (int l = 0; l < k; ++l) s += a[(long)i * lda + l] * b[(long)j * ldb + l];
            c[(long)i * ldc + j] = alpha * s + beta * c[(long)i * ldc + j];
        }
}
extern "C" void qadc_naive_sgemv(enum CBLAS_ORDER, enum CBLAS_TRANSPOSE, int m, int n,
                                 float alpha, const float* a, int lda, const float* x, int,
                                 float beta, float* y, int) {
    std::vector<float> t(m);
    for (int i = 0; i < m; ++i) {
        float s = 0.f;
        for (int l = 0; l < n; ++l) s += a[(long)i * lda + l] * x[l];
        t[i] = alpha * s + beta * y[i];
    }
    std::copy(t.begin(), t.end(), y);
}
#else
extern "C" void scipy_openblas_set_num_threads(int);
#endif

namespace {

// The reference dispatchers (distances.cpp:15-121) reject sq_dim 3 and 6 (96-d configs,
// SURVEY F7); the templates themselves are generic, so instantiate them directly.
dists_mutiple_func harness_multi_func(int sq_dim) {
    switch (sq_dim) {
    case 2: return compute_dists_multiple_blas_cg<2>;
    case 3: return compute_dists_multiple_blas_cg<3>;
    case 6: return compute_dists_multiple_blas_cg<6>;
    case 12: return compute_dists_multiple_blas_cg<12>;
    default: return get_dists_mutiple_function(sq_dim);
    }
}
dists_func harness_single_func(int sq_dim) {
    switch (sq_dim) {
    case 2: return compute_dists_single_simd_cg<2>;
    case 3: return compute_dists_single_simd_cg<3>;
    case 6: return compute_dists_single_simd_cg<6>;
    case 12: return compute_dists_single_simd_cg<12>;
    default: return get_dists_function(sq_dim);
    }
}

struct ref_handle {
    std::unique_ptr<base_db> db;
    std::unique_ptr<scanner_4> scanner;
    bool prepared = false;
};

}  // namespace

extern "C" {

__attribute__((visibility("default"))) int ref_abi_version() { return 1; }

__attribute__((visibility("default"))) void ref_blas_single_thread() {
#ifndef QADC_REF_NAIVE_BLAS
    scipy_openblas_set_num_threads(1);  // README.md:86-94: BLAS must be sequential
#endif
}

// ---- layout (simd_layout.hpp:31-65) -------------------------------------------------
__attribute__((visibility("default"))) long ref_interleaved_size_4(unsigned n, int code_size) {
    return compute_interleaved_size_4(n, code_size, 16);
}
__attribute__((visibility("default"))) void ref_interleave_partition_4(
        std::uint8_t* dst, const std::uint8_t* codes, int code_size, unsigned n) {
    source_partition src{codes, code_size, n};
    interleave_partition_4(dst, src, 16);
}

// ---- the SIMD scan itself (simd_scan.hpp:125-187) -------------------------------------
// qtab: m*16 int8, entry c of sub-quantiser j at qtab[j*16+c] (db_query_4.cpp:65-68).
// The heap is created with capacity r; if push_sentinel, (0,127) is pushed first exactly
// as db_query_4.cpp:276 does. Returns the heap size; keys/vals are the raw heap arrays.
__attribute__((visibility("default"))) int ref_scan_avx_4(
        const std::uint8_t* part, const unsigned* labels, unsigned size, int m,
        const std::int8_t* qtab, int r, int push_sentinel, unsigned* keys_out,
        std::int8_t* vals_out) {
    kv_binheap<unsigned, std::int8_t> bh(r);
    if (push_sentinel) bh.push(0, std::numeric_limits<std::int8_t>::max());
    std::vector<__m128i> q(m);
    for (int j = 0; j < m; ++j)
        q[j] = _mm_loadu_si128(reinterpret_cast<const __m128i*>(qtab + 16 * j));
    if (m == 16) scan_avx_4<16>(part, labels, 0, size, q.data(), bh);
    else if (m == 32) scan_avx_4<32>(part, labels, 0, size, q.data(), bh);
    else return -1;
    std::copy(bh.keys(), bh.keys() + bh.size(), keys_out);
    std::copy(bh.values(), bh.values() + bh.size(), vals_out);
    return bh.size();
}

// Per-vector distances as the reference kernel computes them: run scan_avx_4 with a heap
// that never fills (capacity n+17) seeded with the (0,127) sentinel, so every vector with
// d < 127 is appended; the rest are 127 by elimination.
__attribute__((visibility("default"))) int ref_dump_distances(
        const std::uint8_t* part, unsigned size, int m, const std::int8_t* qtab,
        std::int8_t* out) {
    const int cap = static_cast<int>(size) + 17;
    std::vector<unsigned> keys(cap);
    std::vector<std::int8_t> vals(cap);
    int n = ref_scan_avx_4(part, nullptr, size, m, qtab, cap, 1, keys.data(), vals.data());
    if (n < 0) return n;
    std::fill(out, out + size, static_cast<std::int8_t>(127));
    for (int i = 0; i < n; ++i)
        if (vals[i] != 127) out[keys[i]] = vals[i];
    return 0;
}

// ---- QuantizerMAX<int8_t> (db_query_4.cpp:37-71) --------------------------------------
__attribute__((visibility("default"))) void ref_quantize_tables(
        const float* tables, int m, float qmin, float qmax, std::int8_t* out) {
    QuantizerMAX<std::int8_t> q(qmin, qmax);
    std::vector<__m128i> qt(m);
    q.quantize_tables(tables, qt.data(), m);
    std::memcpy(out, qt.data(), (size_t)m * 16);
}

// ---- float lookup tables (distances.hpp:277-311) --------------------------------------
__attribute__((visibility("default"))) int ref_tables(
        const float* vectors, int count, int dim, int m, const float* codebooks,
        int use_blas_form, float* out) {
    base_pq pq(m, 4, dim, const_cast<float*>(codebooks));
    base_centroids_getter cg(&pq);
    if (use_blas_form) {
        harness_multi_func(pq.sq_dim())(out, cg, vectors, count);
    } else {
        dists_func f = harness_single_func(pq.sq_dim());
        for (int i = 0; i < count; ++i) f(out + (long)i * m * 16, cg, vectors + (long)i * dim);
    }
    return 0;
}

// ---- prefix float ADC (query_common.hpp:59-90) ----------------------------------------
// Row-major codes; returns the heap max after the scan, i.e. what scanner_4 uses as qmax
// (db_query_4.cpp:230-242, :259) when called on one partition prefix.
__attribute__((visibility("default"))) float ref_scan_4_qmax(
        const std::uint8_t* codes, unsigned n, int m, const float* table, int r) {
    kv_binheap<unsigned, float> bh(r);
    bh.push(0, std::numeric_limits<float>::max());
    if (m == 16) scan_4<16>(codes, nullptr, n, table, bh);
    else scan_4<32>(codes, nullptr, n, table, bh);
    return bh.max();
}

// ---- coarse assignment as shipped (neighbors.cpp:30-76, incl. the :64 stride bug) -----
__attribute__((visibility("default"))) void ref_find_k_neighbors(
        int vector_count, int neighbor_count, int dim, int k, const float* vectors,
        const float* neighbors, int* assign) {
    find_k_neighbors(vector_count, neighbor_count, dim, k, vectors, neighbors, assign);
}

__attribute__((visibility("default"))) void ref_residuals(
        const float* vector, int dim, const float* centroids, int* assign, int ma, float* out) {
    substract_vectors_from_unique(vector, dim, centroids, assign, ma, out);
}

// ---- in-memory databases + scanner_4 --------------------------------------------------
// Flat: row-major codes are written straight into flat_db::codes (databases.hpp:78-79).
__attribute__((visibility("default"))) void* ref_flat_create(
        int dim, int m, const float* codebooks, const std::uint8_t* codes, unsigned n) {
    auto h = new ref_handle;
    std::unique_ptr<base_pq> pq(new base_pq(m, 4, dim, const_cast<float*>(codebooks)));
    auto db = new flat_db(std::move(pq));
    db->codes.assign(codes, codes + (size_t)n * (m / 2));
    db->codes_count = n;
    h->db.reset(db);
    return h;
}

// IVF: partition p owns codes[offsets[p]..offsets[p+1]) (row-major) and the same range of
// labels (databases.hpp:179-180).
__attribute__((visibility("default"))) void* ref_ivf_create(
        int dim, int m, const float* codebooks, int K, const float* centroids,
        const std::uint8_t* codes, const unsigned* labels, const long* offsets) {
    auto h = new ref_handle;
    std::unique_ptr<base_pq> pq(new base_pq(m, 4, dim, const_cast<float*>(codebooks)));
    std::unique_ptr<float[]> cents(new float[(size_t)K * dim]);
    std::copy(centroids, centroids + (size_t)K * dim, cents.get());
    auto db = new index_db(std::move(pq), K, std::move(cents));
    const int cs = m / 2;
    for (int p = 0; p < K; ++p) {
        db->partitions[p].assign(codes + offsets[p] * cs, codes + offsets[p + 1] * cs);
        db->labels[p].assign(labels + offsets[p], labels + offsets[p + 1]);
    }
    h->db.reset(db);
    return h;
}

// OPQ: replaces the handle's quantiser by an opq with the same codebooks and the given rotation
// (quantizers.hpp:248-301), before ref_prepare.
__attribute__((visibility("default"))) void ref_set_rotation(void* handle, const float* rotation) {
    auto h = static_cast<ref_handle*>(handle);
    base_pq& old = *h->db->pq;
    h->db->pq.reset(new opq(old.sq_count, old.sq_bits, old.dim, old.centroids_flat.get(),
                            const_cast<float*>(rotation)));
}

// Database files: exactly the calls of flatdb_create.cpp:49-53 (save) and query_common.hpp:321-328
// (load).  Field order comes from the reference's save()/load() members; the bytes of each field
// from shims/cereal (a restatement of cereal 1.2.2's binary archive, see its header).
__attribute__((visibility("default"))) int ref_save_db(void* handle, const char* path) {
    auto h = static_cast<ref_handle*>(handle);
    std::ofstream out_file(path);
    if (!out_file) return -1;
    cereal::BinaryOutputArchive out_archive(out_file);
    out_archive(h->db);
    return out_file ? 0 : -1;
}
__attribute__((visibility("default"))) void* ref_load_db(const char* path) {
    query_args args{};
    args.db_file = path;
    auto h = new ref_handle;
    try {
        h->db = load_database(args);
    } catch (const std::exception& e) {
        std::cerr << e.what() << std::endl;
    }
    if (!h->db) { delete h; return nullptr; }
    return h;
}
// What a loaded database holds: kind 0 flat / 1 index, opq flag, dim, m, bits, partitions.
__attribute__((visibility("default"))) void ref_db_info(void* handle, int* out6) {
    auto h = static_cast<ref_handle*>(handle);
    out6[0] = dynamic_cast<index_db*>(h->db.get()) ? 1 : 0;
    out6[1] = dynamic_cast<opq*>(h->db->pq.get()) ? 1 : 0;
    out6[2] = h->db->pq->dim;
    out6[3] = h->db->pq->sq_count;
    out6[4] = h->db->pq->sq_bits;
    out6[5] = h->db->partition_count();
}

// scanner_4::prepare_database (db_query_4.cpp:210-228). Frees the db's copy of the codes.
__attribute__((visibility("default"))) void ref_prepare(void* handle, float keep) {
    auto h = static_cast<ref_handle*>(handle);
    h->scanner.reset(new scanner_4(keep));
    h->scanner->prepare_database(*h->db);
    h->prepared = true;
}

__attribute__((visibility("default"))) unsigned ref_starts_size(void* handle, int part_i) {
    return static_cast<ref_handle*>(handle)->scanner->starts_sizes[part_i];
}

__attribute__((visibility("default"))) void ref_destroy(void* handle) {
    delete static_cast<ref_handle*>(handle);
}

// scanner_4::query_scan (db_query_4.cpp:245-309) with caller-supplied assign + float
// tables (tables are clamped in place like the reference does). Outputs raw heap arrays.
__attribute__((visibility("default"))) int ref_query_scan(
        void* handle, int* assign, int ma, float* tables, int r, unsigned* keys_out,
        std::int8_t* vals_out) {
    auto h = static_cast<ref_handle*>(handle);
    const int table_dim = h->db->pq->sq_count * 16;
    kv_binheap<unsigned, std::int8_t> bh(r);
    query_metrics metrics;
    h->scanner->query_scan(nullptr, assign, ma, tables, table_dim, bh, metrics);
    std::copy(bh.keys(), bh.keys() + bh.size(), keys_out);
    std::copy(bh.values(), bh.values() + bh.size(), vals_out);
    return bh.size();
}

// The quantiser bounds exactly as scanner_4::query_scan derives them (db_query_4.cpp:250-259):
// query_scan_start over the probed prefixes, then min_element over all ma tables (before the
// negative clamp).
__attribute__((visibility("default"))) void ref_query_bounds(
        void* handle, int* assign, int ma, float* tables, int r, float* qmin, float* qmax) {
    auto h = static_cast<ref_handle*>(handle);
    const int table_dim = h->db->pq->sq_count * 16;
    kv_binheap<unsigned, float> tmp_bh(r);
    h->scanner->query_scan_start(assign, ma, tables, table_dim, tmp_bh);
    *qmin = *std::min_element(tables, tables + (long)ma * table_dim);
    *qmax = tmp_bh.max();
}

// Full per-query path = body of nns_engine::process_query (query_common.hpp:278-307) with
// thread-local scratch, run for nq queries under OpenMP (nthreads; 1 = the reference's own
// sequential loop). use_blas_tables: 0 = what `-b1` does (single_simd when ma==1, blas form
// otherwise), 1 = always the blas form (what nns_engine_batch does).
// keys/vals: nq*r raw heap arrays; sizes: nq. Optional per-query outputs (may be null):
// assign_out nq*ma, tables_out nq*ma*m*16 (after the in-place clamp).
// times_us[0..3] = summed index/rotate/table/scan µs over all queries (per-thread clocks).
__attribute__((visibility("default"))) int ref_search(
        void* handle, const float* queries, int nq, int ma, int r, int nthreads,
        int use_blas_tables, unsigned* keys_out, std::int8_t* vals_out, int* sizes_out,
        int* assign_out, float* tables_out, double* times_us) {
    auto h = static_cast<ref_handle*>(handle);
    base_db& db = *h->db;
    const int dim = db.pq->dim;
    const int m = db.pq->sq_count;
    const int table_dim = m * 16;
    base_centroids_getter cg(db.pq.get());
    dists_func f1 = harness_single_func(db.pq->sq_dim());
    dists_mutiple_func fm = harness_multi_func(db.pq->sq_dim());
    double t_index = 0, t_rot = 0, t_tab = 0, t_scan = 0;
#pragma omp parallel num_threads(nthreads) reduction(+ : t_index, t_rot, t_tab, t_scan)
    {
        std::vector<float> residuals((size_t)ma * dim);
        std::vector<int> assign(ma);
        std::vector<float> tables((size_t)ma * table_dim);
        query_metrics metrics;
#pragma omp for schedule(dynamic)
        for (int qi = 0; qi < nq; ++qi) {
            const float* query = queries + (long)qi * dim;
            const std::uint64_t t0 = ustime();
            db.assign_compute_residuals(query, ma, assign.data(), residuals.data());
            const std::uint64_t t1 = ustime();
            db.pq->rotate_multiple_vectors(residuals.data(), ma);
            const std::uint64_t t2 = ustime();
            if (ma == 1 && !use_blas_tables) f1(tables.data(), cg, residuals.data());
            else fm(tables.data(), cg, residuals.data(), ma);
            const std::uint64_t t3 = ustime();
            kv_binheap<unsigned, std::int8_t> bh(r);
            h->scanner->query_scan(residuals.data(), assign.data(), ma, tables.data(),
                                   table_dim, bh, metrics);
            const std::uint64_t t4 = ustime();
            t_index += t1 - t0; t_rot += t2 - t1; t_tab += t3 - t2; t_scan += t4 - t3;
            sizes_out[qi] = bh.size();
            std::copy(bh.keys(), bh.keys() + bh.size(), keys_out + (long)qi * r);
            std::copy(bh.values(), bh.values() + bh.size(), vals_out + (long)qi * r);
            if (assign_out) std::copy(assign.begin(), assign.end(), assign_out + (long)qi * ma);
            if (tables_out)
                std::copy(tables.begin(), tables.end(), tables_out + (long)qi * ma * table_dim);
        }
    }
    if (times_us) { times_us[0] = t_index; times_us[1] = t_rot; times_us[2] = t_tab; times_us[3] = t_scan; }
    return 0;
}

// PQ encoder as shipped (quantizers.hpp:222-245): floats -> row-major 4-bit codes.
// NOTE: rotates/modifies `vectors` in place for OPQ; base_pq leaves them untouched.
__attribute__((visibility("default"))) void ref_encode(
        int dim, int m, const float* codebooks, float* vectors, int count, std::uint8_t* codes) {
    base_pq pq(m, 4, dim, const_cast<float*>(codebooks));
    pq.encode_multiple_vectors(vectors, codes, count);
}

// Same for any sq_bits the reference packs (4 or 8 here): base_pq(m, bits, ...).
__attribute__((visibility("default"))) void ref_encode_bits(
        int dim, int m, int bits, const float* codebooks, float* vectors, int count, std::uint8_t* codes) {
    base_pq pq(m, bits, dim, const_cast<float*>(codebooks));
    pq.encode_multiple_vectors(vectors, codes, count);
}

// index_db::add_vectors as shipped (databases.hpp:270-298), optionally with an opq (residual -> rotate -> encode,
// quantizers.hpp:222-224, :289-301): per inserted vector the cell it went to and the code that was stored.
// NOTE: modifies `vectors` in place (residuals, rotation) like the reference does.
__attribute__((visibility("default"))) void ref_index_add_vectors(
        int dim, int m, const float* codebooks, const float* rotation, int K, const float* centroids,
        float* vectors, unsigned count, int* out_assign, std::uint8_t* out_codes) {
    std::unique_ptr<base_pq> pq;
    if (rotation) pq.reset(new opq(m, 4, dim, const_cast<float*>(codebooks), const_cast<float*>(rotation)));
    else pq.reset(new base_pq(m, 4, dim, const_cast<float*>(codebooks)));
    std::unique_ptr<float[]> cents(new float[(size_t)K * dim]);
    std::copy(centroids, centroids + (size_t)K * dim, cents.get());
    index_db db(std::move(pq), K, std::move(cents));
    db.add_vectors(vectors, count, 0, 1);
    const int cs = m / 2;
    for (int p = 0; p < K; ++p)
        for (size_t i = 0; i < db.labels[p].size(); ++i) {
            const unsigned v = db.labels[p][i];
            out_assign[v] = p;
            std::copy(db.partitions[p].begin() + i * cs, db.partitions[p].begin() + (i + 1) * cs, out_codes + (size_t)v * cs);
        }
}

}  // extern "C"
