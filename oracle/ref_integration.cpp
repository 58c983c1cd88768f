// oracle/ref_integration.cpp — TEST INFRASTRUCTURE ONLY.
//
// Proves the integration claim of INTEGRATION.md: the UNMODIFIED reference translation unit
// (/root/reference/db_query_4.cpp, compiled where it lies, main() renamed) plus the adaptor header a maintainer
// would add (quick-adc_b200/host/reference_adaptor.hpp, written against the reference's own types) and
// libqadc_b200.so.  The reference's own process_queries<Engine, Bh, Metrics> (query_common.hpp:330-368) is
// instantiated twice on the same in-memory database: with nns_engine_batch<scanner_4> (the reference) and with
// nns_engine_gpu (the B200 path).  A thin recording wrapper, identical for both, copies every query's heap
// out of the loop so that tests/test_gpu_integration.py can compare them (tie-class rule, SURVEY §8c Stage R).
#define main qadc_ref_main
#include "db_query_4.cpp"
#undef main

#include "../quick-adc_b200/host/reference_adaptor.hpp"

extern "C" void scipy_openblas_set_num_threads(int) __attribute__((weak));

namespace {

// Engine concept wrapper: forwards to the wrapped engine and records the caller's heap arrays (heap order).
template <typename Engine>
struct recording_engine {
    Engine& inner;
    int r;
    unsigned* keys; std::int8_t* vals; int* sizes;
    void prepare_database() { inner.prepare_database(); }
    template <typename DistType, typename MetricsType>
    void process_query(const int query_i, const float* queries, const int count, kv_binheap<unsigned, DistType>& bh,
                       MetricsType& metrics) {
        inner.process_query(query_i, queries, count, bh, metrics);
        sizes[query_i] = bh.size();
        std::copy(bh.keys(), bh.keys() + bh.size(), keys + static_cast<long>(query_i) * r);
        std::copy(bh.values(), bh.values() + bh.size(), vals + static_cast<long>(query_i) * r);
    }
};

std::unique_ptr<base_db> make_db(int dim, int m, const float* codebooks, const float* rotation, int K,
                                 const float* centroids, const std::uint8_t* codes, const unsigned* labels,
                                 const long* offsets) {
    std::unique_ptr<base_pq> pq;
    if (rotation) pq.reset(new opq(m, 4, dim, const_cast<float*>(codebooks), const_cast<float*>(rotation)));
    else pq.reset(new base_pq(m, 4, dim, const_cast<float*>(codebooks)));
    const int cs = m / 2;
    if (K == 0) {
        auto db = new flat_db(std::move(pq));
        db->codes.assign(codes + offsets[0] * cs, codes + offsets[1] * cs);
        db->codes_count = static_cast<unsigned>(offsets[1] - offsets[0]);
        return std::unique_ptr<base_db>(db);
    }
    std::unique_ptr<float[]> cents(new float[static_cast<size_t>(K) * dim]);
    std::copy(centroids, centroids + static_cast<size_t>(K) * dim, cents.get());
    auto db = new index_db(std::move(pq), K, std::move(cents));
    for (int p = 0; p < K; ++p) {
        db->partitions[p].assign(codes + offsets[p] * cs, codes + offsets[p + 1] * cs);
        db->labels[p].assign(labels + offsets[p], labels + offsets[p + 1]);
    }
    return std::unique_ptr<base_db>(db);
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int refint_run(
        int dim, int m, const float* codebooks, const float* rotation, int K, const float* centroids,
        const std::uint8_t* codes, const unsigned* labels, const long* offsets, float keep,
        const char* query_file, const char* groundtruth_file, int r, int ma, int batch, int use_gpu,
        const int* devices, int n_devices,
        unsigned* out_keys, std::int8_t* out_vals, int* out_sizes, double* out_recall, double* out_metrics4) {
    if (scipy_openblas_set_num_threads) scipy_openblas_set_num_threads(1);
    std::unique_ptr<base_db> db = make_db(dim, m, codebooks, rotation, K, centroids, codes, labels, offsets);
    query_args args;
    args.db_file = "";
    args.query_file = query_file;
    args.groundtruth_file = groundtruth_file;
    args.r = r;
    args.ma = ma;
    query_metrics total_metrics;
    double total_recall = 0;
    if (use_gpu) {
        std::vector<int> dev(devices, devices + n_devices);
        std::unique_ptr<scanner_gpu_4> scanner(new scanner_gpu_4(keep, dev));
        nns_engine_gpu engine(std::move(scanner), *db, ma, r, batch);
        recording_engine<nns_engine_gpu> rec{engine, r, out_keys, out_vals, out_sizes};
        process_queries<recording_engine<nns_engine_gpu>, scanner_gpu_4::BhType>(args, *db, rec, total_metrics, total_recall);
    } else {
        std::unique_ptr<base_centroids_getter> cg(new base_centroids_getter(db->pq.get()));
        std::unique_ptr<scanner_4> scanner(new scanner_4(keep));
        nns_engine_batch<scanner_4> engine(std::move(scanner), std::move(cg), *db, ma, batch);
        recording_engine<nns_engine_batch<scanner_4>> rec{engine, r, out_keys, out_vals, out_sizes};
        process_queries<recording_engine<nns_engine_batch<scanner_4>>, scanner_4::BhType>(args, *db, rec, total_metrics, total_recall);
    }
    *out_recall = total_recall;
    out_metrics4[0] = total_metrics.index_us; out_metrics4[1] = total_metrics.rotate_us;
    out_metrics4[2] = total_metrics.table_us; out_metrics4[3] = total_metrics.scan_us;
    return 0;
}
