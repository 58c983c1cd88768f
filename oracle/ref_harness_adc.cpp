// oracle/ref_harness_adc.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// The plain ADC query tool of the reference (db_query.cpp: scanner_simple + scan_standard /
// scan_4, query_common.hpp:59-147) behind a small C ABI, compiled from the UNMODIFIED sources
// where they lie.  db_query.cpp and db_query_4.cpp define the same global names (main, usage,
// cmdargs, parse_args, process_queries) and query_common.hpp defines non-inline functions, so
// this is a second shared library (oracle/_ref/libqadc_ref_adc.so) next to libqadc_ref.so.
// The only restated logic is the per-query body of nns_engine::process_query
// (query_common.hpp:278-307), as in ref_harness.cpp.
#define main qadc_ref_adc_main
#include "db_query.cpp"
#undef main

#include <cstring>
#include <vector>

extern "C" void scipy_openblas_set_num_threads(int);

namespace {
// same instantiations as ref_harness.cpp: the dispatchers reject sq_dim 3 and 6
dists_func adc_single_func(int sq_dim) {
    switch (sq_dim) {
    case 2: return compute_dists_single_simd_cg<2>;
    case 3: return compute_dists_single_simd_cg<3>;
    case 6: return compute_dists_single_simd_cg<6>;
    case 12: return compute_dists_single_simd_cg<12>;
    default: return get_dists_function(sq_dim);
    }
}
dists_mutiple_func adc_multi_func(int sq_dim) {
    switch (sq_dim) {
    case 2: return compute_dists_multiple_blas_cg<2>;
    case 3: return compute_dists_multiple_blas_cg<3>;
    case 6: return compute_dists_multiple_blas_cg<6>;
    case 12: return compute_dists_multiple_blas_cg<12>;
    default: return get_dists_mutiple_function(sq_dim);
    }
}
struct adc_handle {
    std::unique_ptr<base_db> db;
    scanner_simple scanner;
};
}  // namespace

extern "C" {

// codes: row-major, partition p = [offsets[p], offsets[p+1]); K == 0: flat database (one partition).
__attribute__((visibility("default"))) void* refadc_create(
        int dim, int m, int bits, const float* codebooks, const float* rotation, int K,
        const float* centroids, const std::uint8_t* codes, const unsigned* labels, const long* offsets) {
    scipy_openblas_set_num_threads(1);   // README.md:86-94
    auto h = new adc_handle;
    std::unique_ptr<base_pq> pq;
    if (rotation) pq.reset(new opq(m, bits, dim, const_cast<float*>(codebooks), const_cast<float*>(rotation)));
    else pq.reset(new base_pq(m, bits, dim, const_cast<float*>(codebooks)));
    const long cs = (long)m * bits / 8;
    if (K == 0) {
        auto db = new flat_db(std::move(pq));
        db->codes.assign(codes + offsets[0] * cs, codes + offsets[1] * cs);
        db->codes_count = (unsigned)(offsets[1] - offsets[0]);
        h->db.reset(db);
    } else {
        std::unique_ptr<float[]> cents(new float[(size_t)K * dim]);
        std::copy(centroids, centroids + (size_t)K * dim, cents.get());
        auto db = new index_db(std::move(pq), K, std::move(cents));
        for (int p = 0; p < K; ++p) {
            db->partitions[p].assign(codes + offsets[p] * cs, codes + offsets[p + 1] * cs);
            db->labels[p].assign(labels + offsets[p], labels + offsets[p + 1]);
        }
        h->db.reset(db);
    }
    h->scanner.prepare_database(*h->db);   // get_scan_func: exits on unsupported (nsq, bits)
    return h;
}

__attribute__((visibility("default"))) void refadc_destroy(void* handle) { delete static_cast<adc_handle*>(handle); }

// Per query: assign + residuals, rotation, float tables (single_simd when ma == 1 like `-b1`,
// else the blas form), scanner_simple::query_scan; the heap is returned sorted by distance
// (kv_binheap::sort, values only).  ids/dists: nq*r.
__attribute__((visibility("default"))) int refadc_search(
        void* handle, const float* queries, int nq, int ma, int r, unsigned* ids_out, float* dists_out) {
    auto h = static_cast<adc_handle*>(handle);
    base_db& db = *h->db;
    const int dim = db.pq->dim;
    const int table_dim = db.pq->sq_count * db.pq->sq_centroid_count();
    base_centroids_getter cg(db.pq.get());
    dists_func f1 = adc_single_func(db.pq->sq_dim());
    dists_mutiple_func fm = adc_multi_func(db.pq->sq_dim());
    std::vector<float> residuals((size_t)ma * dim), tables((size_t)ma * table_dim);
    std::vector<int> assign(ma);
    query_metrics metrics;
    for (int qi = 0; qi < nq; ++qi) {
        const float* query = queries + (long)qi * dim;
        db.assign_compute_residuals(query, ma, assign.data(), residuals.data());
        db.pq->rotate_multiple_vectors(residuals.data(), ma);
        if (ma == 1) f1(tables.data(), cg, residuals.data());
        else fm(tables.data(), cg, residuals.data(), ma);
        kv_binheap<unsigned, float> bh(r);
        h->scanner.query_scan(residuals.data(), assign.data(), ma, tables.data(), table_dim, bh, metrics);
        bh.sort(ids_out + (long)qi * r, dists_out + (long)qi * r);
    }
    return 0;
}

}  // extern "C"
