// Engine side of the drop-in (query_common.hpp:21-56 query_metrics, :149-243 nns_engine_batch,
// :330-368 process_queries, recall.hpp:45-54).  nns_engine_gpu satisfies the reference's Engine
// concept — prepare_database(); process_query(query_i, queries, count, bh, metrics) — and does
// what nns_engine_batch does at `query_i % batch == 0`: one call for the whole batch, here
// qadc_search over the C ABI; every call then copies query_i's result into the caller's heap.
#ifndef QADC_HOST_QUERY_COMMON_HPP_
#define QADC_HOST_QUERY_COMMON_HPP_

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <iostream>
#include <limits>
#include <memory>
#include <vector>

#include "../../include/qadc_b200.h"
#include "binheap.hpp"
#include "databases.hpp"
#include "vector_io.hpp"

struct query_metrics {
    std::uint64_t index_us = 0, rotate_us = 0, table_us = 0, scan_us = 0;
    static constexpr const char* header_string = "index_us,rotate_us,table_us,scan_us";
    query_metrics& operator+=(const query_metrics& r) {
        index_us += r.index_us; rotate_us += r.rotate_us; table_us += r.table_us; scan_us += r.scan_us;
        return *this;
    }
    query_metrics& operator/=(const int f) {
        index_us /= f; rotate_us /= f; table_us /= f; scan_us /= f;
        return *this;
    }
};
inline std::ostream& operator<<(std::ostream& os, const query_metrics& m) {
    return os << m.index_us << "," << m.rotate_us << "," << m.table_us << "," << m.scan_us;
}

struct query_args {
    const char* db_file;
    const char* query_file;
    const char* groundtruth_file;
    int r;
    int ma;
};

[[noreturn]] inline void qadc_die(qadc_ctx* ctx, const char* what) {
    std::cerr << what << ": " << qadc_last_error(ctx) << std::endl;   // the reference: cerr + exit(1)
    std::exit(1);
}

[[noreturn]] inline void qadc_multi_die(qadc_multi* mm, const char* what) {
    std::cerr << what << ": " << qadc_multi_last_error(mm) << std::endl;
    std::exit(1);
}

// Scanner concept (db_query_4.cpp:73-310): owns the device copy of the database — on one GPU (a qadc_ctx) or
// sharded over several GPUs of the box by this one process (a qadc_multi: contiguous runs of a flat database /
// whole inverted lists per GPU, NCCL all-gather of the per-shard top-r lists, merge on the first device).
struct scanner_gpu_4 {
    typedef kv_binheap<unsigned, std::int8_t> BhType;
    qadc_ctx* ctx = nullptr;        // one device
    qadc_multi* multi = nullptr;    // several devices
    float keep;
    explicit scanner_gpu_4(float keep_, int device = 0) : keep(keep_) {
        if (qadc_create(device, nullptr, &ctx) != QADC_OK) qadc_die(nullptr, "qadc_create");
    }
    scanner_gpu_4(float keep_, const std::vector<int>& devices) : keep(keep_) {
        if (devices.size() == 1) {
            if (qadc_create(devices[0], nullptr, &ctx) != QADC_OK) qadc_die(nullptr, "qadc_create");
        } else if (qadc_multi_create(devices.data(), static_cast<int>(devices.size()), &multi) != QADC_OK) {
            qadc_multi_die(nullptr, "qadc_multi_create");
        }
    }
    ~scanner_gpu_4() { qadc_destroy(ctx); qadc_multi_destroy(multi); }
    scanner_gpu_4(const scanner_gpu_4&) = delete;

    // scanner_4::prepare_database (db_query_4.cpp:210-228): copies every partition to the device(s)
    // (re-laid out there) and frees the database's own copy (:190).
    void prepare_database(base_db& db) {
        const base_pq& pq = *db.pq;
        const int parts = db.partition_count();
        std::vector<std::uint32_t> sizes(parts);
        std::vector<const std::uint8_t*> part_codes(parts, nullptr);
        std::vector<const std::uint32_t*> part_labels(parts, nullptr);
        const std::uint8_t* codes;
        unsigned* labels;
        unsigned size;
        bool has_labels = false, any = false;
        for (int p = 0; p < parts; ++p) {
            db.get_partition(p, codes, labels, size);
            sizes[p] = size;
            if (size == 0) { std::cerr << "Warning: Partition " << p << " is empty" << std::endl; continue; }
            if (any && (labels != nullptr) != has_labels) {
                std::cerr << "Cannot prepare database. Some partitions have labels and some have not" << std::endl;
                std::exit(1);
            }
            any = true;
            has_labels = labels != nullptr;
            part_codes[p] = codes;
            part_labels[p] = labels;
        }
        if (multi) {
            if (qadc_multi_set_pq(multi, pq.dim, pq.sq_count, pq.sq_bits, pq.centroids_flat.data(), pq.rotation_ptr())) qadc_multi_die(multi, "qadc_multi_set_pq");
            if (db.coarse_centroids() && qadc_multi_set_coarse(multi, parts, db.coarse_centroids())) qadc_multi_die(multi, "qadc_multi_set_coarse");
            if (qadc_multi_load(multi, parts, sizes.data(), part_codes.data(), has_labels ? part_labels.data() : nullptr, keep))
                qadc_multi_die(multi, "qadc_multi_load");
        } else {
            if (qadc_set_pq(ctx, pq.dim, pq.sq_count, pq.sq_bits, pq.centroids_flat.data(), pq.rotation_ptr())) qadc_die(ctx, "qadc_set_pq");
            if (db.coarse_centroids() && qadc_set_coarse(ctx, parts, db.coarse_centroids())) qadc_die(ctx, "qadc_set_coarse");
            if (qadc_begin_database(ctx, parts, sizes.data(), has_labels)) qadc_die(ctx, "qadc_begin_database");
            // one copy + one re-layout launch per run of partitions (65 536 inverted lists: not 65 536 launches)
            if (qadc_upload_partitions(ctx, part_codes.data(), has_labels ? part_labels.data() : nullptr)) qadc_die(ctx, "qadc_upload_partitions");
            if (qadc_finalize(ctx, keep)) qadc_die(ctx, "qadc_finalize");
        }
        for (int p = 0; p < parts; ++p) db.free_partition(p);
    }

    // one batch: nns_engine_batch::batch_process_queries + per-query scanner_4::query_scan
    void search(const float* queries, int nq, int ma, int r, std::uint32_t* ids, std::int8_t* dists, std::int32_t* counts,
                qadc_metrics* m) {
        if (multi) {
            if (qadc_multi_search(multi, queries, nq, ma, r, ids, dists, counts, m)) qadc_multi_die(multi, "qadc_multi_search");
        } else if (qadc_search(ctx, queries, nq, ma, r, ids, dists, counts, m)) {
            qadc_die(ctx, "qadc_search");
        }
    }
};

// Engine concept (query_common.hpp:149-243).
struct nns_engine_gpu {
    base_db& db_;
    std::unique_ptr<scanner_gpu_4> scanner_;
    int ma_, r_, batch_count_;
    std::vector<std::uint32_t> ids_;
    std::vector<std::int8_t> dists_;
    std::vector<std::int32_t> counts_;
    int batch_first_ = -1, batch_size_ = 0;

    nns_engine_gpu(std::unique_ptr<scanner_gpu_4>&& scanner, base_db& db, int ma, int r, int batch_count)
        : db_(db), scanner_(std::move(scanner)), ma_(ma), r_(r), batch_count_(batch_count <= 0 ? 1 << 14 : batch_count) {
        std::cerr << "NNS Engine Batch size: " << batch_count_ << " queries" << std::endl;
    }
    void prepare_database() { scanner_->prepare_database(db_); }

    template <typename DistType, typename MetricsType>
    void process_query(const int query_i, const float* queries, const int count, kv_binheap<unsigned, DistType>& bh,
                       MetricsType& metrics) {
        metrics = MetricsType();
        if (query_i % batch_count_ == 0 || query_i < batch_first_ || query_i >= batch_first_ + batch_size_) {
            batch_first_ = (query_i / batch_count_) * batch_count_;
            batch_size_ = std::min(batch_count_, count - batch_first_);
            ids_.resize(static_cast<size_t>(batch_size_) * r_);
            dists_.resize(ids_.size());
            counts_.resize(batch_size_);
            qadc_metrics m;
            const int dim = db_.pq->dim;
            scanner_->search(queries + static_cast<long>(batch_first_) * dim, batch_size_, ma_, r_, ids_.data(), dists_.data(),
                             counts_.data(), &m);
            // like nns_engine_batch, the batch-level phases are booked on the first query of the batch
            metrics.index_us = static_cast<std::uint64_t>(m.index_us);
            metrics.rotate_us = static_cast<std::uint64_t>(m.rotate_us);
            metrics.table_us = static_cast<std::uint64_t>(m.table_us);
            metrics.scan_us = static_cast<std::uint64_t>(m.scan_us + m.h2d_us + m.d2h_us);
        }
        const int b = query_i - batch_first_;
        // result i -> the caller's heap: sentinel first (db_query_4.cpp:276), then the canonical list
        bh.push(0, static_cast<DistType>(127));
        for (int i = 0; i < counts_[b]; ++i)
            bh.push(ids_[static_cast<size_t>(b) * r_ + i], static_cast<DistType>(dists_[static_cast<size_t>(b) * r_ + i]));
    }
};

inline std::uint64_t ustime() {   // common.hpp:15-20
    return static_cast<std::uint64_t>(std::chrono::duration_cast<std::chrono::microseconds>(
        std::chrono::steady_clock::now().time_since_epoch()).count());
}

// ---- plain ADC (db_query.cpp): scanner_simple + engine on the GPU float scan ------------------
// Scanner concept of db_query.cpp:17-46: the database goes to the device row-major as it is.
struct scanner_gpu_simple {
    typedef kv_binheap<unsigned, float> BhType;
    qadc_ctx* ctx = nullptr;
    explicit scanner_gpu_simple(int device = 0) {
        if (qadc_create(device, nullptr, &ctx) != QADC_OK) qadc_die(nullptr, "qadc_create");
    }
    ~scanner_gpu_simple() { qadc_destroy(ctx); }
    scanner_gpu_simple(const scanner_gpu_simple&) = delete;

    void prepare_database(base_db& db) {
        const base_pq& pq = *db.pq;
        // get_scan_func (query_common.hpp:122-147) rejects other (nsq, bits) pairs with the same words
        if (qadc_set_pq(ctx, pq.dim, pq.sq_count, pq.sq_bits, pq.centroids_flat.data(), pq.rotation_ptr())) qadc_die(ctx, "qadc_set_pq");
        const int parts = db.partition_count();
        if (db.coarse_centroids() && qadc_set_coarse(ctx, parts, db.coarse_centroids())) qadc_die(ctx, "qadc_set_coarse");
        const size_t cs = pq.code_size();
        std::vector<std::uint64_t> offsets(parts + 1, 0);
        std::vector<std::uint8_t> codes_all;
        std::vector<std::uint32_t> labels_all;
        const std::uint8_t* codes;
        unsigned* labels;
        unsigned size;
        for (int p = 0; p < parts; ++p) {
            db.get_partition(p, codes, labels, size);
            offsets[p + 1] = offsets[p] + size;
            codes_all.insert(codes_all.end(), codes, codes + static_cast<size_t>(size) * cs);
            if (labels) labels_all.insert(labels_all.end(), labels, labels + size);
            db.free_partition(p);
        }
        if (qadc_adc_load(ctx, parts, offsets.data(), codes_all.data(), db.coarse_centroids() ? labels_all.data() : nullptr))
            qadc_die(ctx, "qadc_adc_load");
    }
};

struct nns_engine_gpu_adc {
    base_db& db_;
    std::unique_ptr<scanner_gpu_simple> scanner_;
    int ma_, r_, batch_count_;
    std::vector<std::uint32_t> ids_;
    std::vector<float> dists_;
    std::vector<std::int32_t> counts_;
    int batch_first_ = -1, batch_size_ = 0;

    nns_engine_gpu_adc(std::unique_ptr<scanner_gpu_simple>&& scanner, base_db& db, int ma, int r, int batch_count)
        : db_(db), scanner_(std::move(scanner)), ma_(ma), r_(r), batch_count_(batch_count <= 0 ? 1 << 14 : batch_count) {
        std::cerr << "NNS Engine Batch size: " << batch_count_ << " queries" << std::endl;
    }
    void prepare_database() { scanner_->prepare_database(db_); }

    template <typename DistType, typename MetricsType>
    void process_query(const int query_i, const float* queries, const int count, kv_binheap<unsigned, DistType>& bh,
                       MetricsType& metrics) {
        metrics = MetricsType();
        if (query_i % batch_count_ == 0 || query_i < batch_first_ || query_i >= batch_first_ + batch_size_) {
            batch_first_ = (query_i / batch_count_) * batch_count_;
            batch_size_ = std::min(batch_count_, count - batch_first_);
            ids_.resize(static_cast<size_t>(batch_size_) * r_);
            dists_.resize(ids_.size());
            counts_.resize(batch_size_);
            const std::uint64_t t0 = ustime();
            if (qadc_adc_search(scanner_->ctx, queries + static_cast<long>(batch_first_) * db_.pq->dim, batch_size_, ma_, r_,
                                ids_.data(), dists_.data(), counts_.data()))
                qadc_die(scanner_->ctx, "qadc_adc_search");
            metrics.scan_us = ustime() - t0;   // whole batch (assignment, tables, scan, copies), booked on its first query
        }
        const int b = query_i - batch_first_;
        // scanner_simple::query_scan (db_query.cpp:27-30) pre-fills the heap, then the scan replaces entries
        for (int t = 0; t < bh.capacity(); ++t) bh.push(0, std::numeric_limits<DistType>::max() - t);
        for (int i = 0; i < counts_[b]; ++i)
            bh.push(ids_[static_cast<size_t>(b) * r_ + i], static_cast<DistType>(dists_[static_cast<size_t>(b) * r_ + i]));
    }
};

// recall.hpp:45-54 with t = 1 (query_common.hpp:342): 1 iff the first ground-truth id is returned.
struct recall_file {
    vectors_owner<int> groundtruth;
    explicit recall_file(const char* filename) { groundtruth = load_vecs_as<int, std::int32_t>(filename); }
    template <typename It>
    int check_labels(const int query_i, It first, It last, const int t) const {
        const int* gt = groundtruth.get(query_i);
        for (int i = 0; i < t; ++i)
            if (std::find(first, last, static_cast<unsigned>(gt[i])) == last) return 0;
        return 1;
    }
};

// query_common.hpp:330-368
template <typename EngineType, typename BhType, typename MetricsType, typename DumpT = std::int8_t>
void process_queries(query_args& args, base_db& db, EngineType& engine, MetricsType& total_metrics, double& total_recall,
                     std::vector<unsigned>* dump_keys = nullptr, std::vector<DumpT>* dump_vals = nullptr,
                     int max_queries = -1) {
    vectors_owner<float> queries = load_vectors_by_extension(args.query_file);
    if (max_queries > 0) queries.count = max_queries;
    if (queries.dimension != db.pq->dim) {
        std::cerr << "Query dimension " << queries.dimension << " != database dimension " << db.pq->dim << std::endl;
        std::exit(1);
    }
    recall_file rec_file(args.groundtruth_file);
    const int t = 1;
    MetricsType metrics;
    engine.prepare_database();
    const float* queries_buffer = queries.get(0);
    for (int query_i = 0; query_i < queries.count; ++query_i) {
        BhType bh(args.r);
        engine.process_query(query_i, queries_buffer, static_cast<int>(queries.count), bh, metrics);
        if (bh.size() != args.r) std::cerr << " WARNING: Binheap not full" << std::endl;
        total_recall += rec_file.check_labels(query_i, bh.keys(), bh.keys() + bh.size(), t);
        total_metrics += metrics;
        if (dump_keys) {
            std::vector<unsigned> k(args.r, 0);
            std::vector<DumpT> v(args.r, std::numeric_limits<DumpT>::max());
            bh.sort(k.data(), v.data());
            dump_keys->insert(dump_keys->end(), k.begin(), k.end());
            dump_vals->insert(dump_vals->end(), v.begin(), v.end());
        }
    }
    total_metrics /= static_cast<int>(queries.count);
    total_recall /= queries.count;
}

#endif
