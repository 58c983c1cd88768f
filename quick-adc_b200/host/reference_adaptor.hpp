// reference_adaptor.hpp — the file a maintainer of technicolor-research/quick-adc drops into the reference
// tree to run its search path on B200s: include it in db_query_4.cpp after the reference's own headers
// (binheap.hpp, databases.hpp, quantizers.hpp, query_common.hpp) and link libqadc_b200.so.
//
// It is written against the REFERENCE's types, not against this repository's host mirror:
//   scanner_gpu_4   Scanner concept (replaces scanner_4, db_query_4.cpp:73-310)
//   nns_engine_gpu  Engine concept  (replaces nns_engine_batch<scanner_4>, query_common.hpp:149-243)
// so that the reference's own process_queries<Engine, Bh, Metrics> (query_common.hpp:330-368), its heap, its
// recall check and its CSV line run unchanged.  oracle/ref_integration.cpp compiles exactly this header
// against /root/reference and tests/test_gpu_integration.py compares it with nns_engine_batch<scanner_4>.
#ifndef QADC_REFERENCE_ADAPTOR_HPP_
#define QADC_REFERENCE_ADAPTOR_HPP_

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <memory>
#include <vector>

#include "qadc_b200.h"

struct scanner_gpu_4 {
    typedef kv_binheap<unsigned, std::int8_t> BhType;
    qadc_ctx* ctx = nullptr;        // one device
    qadc_multi* multi = nullptr;    // several devices (database sharded by this process)
    float keep;

    [[noreturn]] void die(const char* what) const {   // the reference's error behaviour: message + exit(1)
        std::cerr << what << ": " << (multi ? qadc_multi_last_error(multi) : qadc_last_error(ctx)) << std::endl;
        std::exit(1);
    }
    explicit scanner_gpu_4(float keep_, const std::vector<int>& devices = std::vector<int>(1, 0)) : keep(keep_) {
        const int rc = devices.size() == 1 ? qadc_create(devices[0], nullptr, &ctx)
                                           : qadc_multi_create(devices.data(), static_cast<int>(devices.size()), &multi);
        if (rc) die("qadc_create");
    }
    ~scanner_gpu_4() { qadc_destroy(ctx); qadc_multi_destroy(multi); }
    scanner_gpu_4(const scanner_gpu_4&) = delete;

    // == scanner_4::prepare_database (db_query_4.cpp:210-228): the partitions go to the device(s), the database's
    // own copy is freed (:190)
    void prepare_database(base_db& db) {
        base_pq& pq = *db.pq;
        opq* o = dynamic_cast<opq*>(&pq);
        index_db* x = dynamic_cast<index_db*>(&db);
        const float* rot = o ? o->rotation.get() : nullptr;
        const int P = db.partition_count();
        std::vector<std::uint32_t> sizes(P);
        std::vector<const std::uint8_t*> codes(P, nullptr);
        std::vector<const std::uint32_t*> labels(P, nullptr);
        bool has_labels = false;
        for (int p = 0; p < P; ++p) {
            const std::uint8_t* c; unsigned* l; unsigned n;
            db.get_partition(p, c, l, n);
            sizes[p] = n;
            if (n == 0) { std::cerr << "Warning: Partition " << p << " is empty" << std::endl; continue; }
            codes[p] = c; labels[p] = l;
            has_labels = has_labels || l != nullptr;
        }
        if (multi) {
            if (qadc_multi_set_pq(multi, pq.dim, pq.sq_count, pq.sq_bits, pq.centroids_flat.get(), rot)) die("qadc_multi_set_pq");
            if (x && qadc_multi_set_coarse(multi, x->part_count, x->centroids.get())) die("qadc_multi_set_coarse");
            if (qadc_multi_load(multi, P, sizes.data(), codes.data(), has_labels ? labels.data() : nullptr, keep)) die("qadc_multi_load");
        } else {
            if (qadc_set_pq(ctx, pq.dim, pq.sq_count, pq.sq_bits, pq.centroids_flat.get(), rot)) die("qadc_set_pq");
            if (x && qadc_set_coarse(ctx, x->part_count, x->centroids.get())) die("qadc_set_coarse");
            if (qadc_begin_database(ctx, P, sizes.data(), has_labels)) die("qadc_begin_database");
            if (qadc_upload_partitions(ctx, codes.data(), has_labels ? labels.data() : nullptr)) die("qadc_upload_partitions");
            if (qadc_finalize(ctx, keep)) die("qadc_finalize");
        }
        for (int p = 0; p < P; ++p) db.free_partition(p);
    }

    void search(const float* queries, int nq, int ma, int r, std::uint32_t* ids, std::int8_t* dists, std::int32_t* counts,
                qadc_metrics* m) {
        const int rc = multi ? qadc_multi_search(multi, queries, nq, ma, r, ids, dists, counts, m)
                             : qadc_search(ctx, queries, nq, ma, r, ids, dists, counts, m);
        if (rc) die("qadc_search");
    }
};

struct nns_engine_gpu {
    base_db& db_;
    std::unique_ptr<scanner_gpu_4> scanner_;
    int ma_, r_, batch_count_, dim_;
    std::vector<std::uint32_t> ids_;
    std::vector<std::int8_t> dists_;
    std::vector<std::int32_t> counts_;

    // batch_count <= 0: 16 384 queries per device call (the reference sizes its batches by a 1 GiB table buffer,
    // query_common.hpp:147, :171-175; here the tables never leave the device)
    nns_engine_gpu(std::unique_ptr<scanner_gpu_4>&& scanner, base_db& db, int ma, int r, int batch_count = -1)
        : db_(db), scanner_(std::move(scanner)), ma_(ma), r_(r), batch_count_(batch_count <= 0 ? 1 << 14 : batch_count),
          dim_(db.pq->dim) {
        std::cerr << "NNS Engine Batch size: " << batch_count_ << " queries" << std::endl;
    }
    void prepare_database() { scanner_->prepare_database(db_); }

    template <typename DistType, typename MetricsType>
    void process_query(const int query_i, const float* queries, const int count, kv_binheap<unsigned, DistType>& bh,
                       MetricsType& metrics) {
        const int b = query_i % batch_count_;
        metrics.index_us = metrics.rotate_us = metrics.table_us = metrics.scan_us = 0;
        if (b == 0) {   // same place as nns_engine_batch::batch_process_queries (query_common.hpp:225-227)
            const int n = std::min(batch_count_, count - query_i);
            ids_.resize(static_cast<std::size_t>(n) * r_);
            dists_.resize(ids_.size());
            counts_.resize(n);
            qadc_metrics m;
            scanner_->search(queries + static_cast<long>(query_i) * dim_, n, ma_, r_, ids_.data(), dists_.data(), counts_.data(), &m);
            metrics.index_us = static_cast<std::uint64_t>(m.index_us);
            metrics.rotate_us = static_cast<std::uint64_t>(m.rotate_us);
            metrics.table_us = static_cast<std::uint64_t>(m.table_us);
            metrics.scan_us = static_cast<std::uint64_t>(m.scan_us + m.h2d_us + m.d2h_us);
        }
        bh.push(0, static_cast<DistType>(127));   // the reference's sentinel (db_query_4.cpp:276)
        for (int i = 0; i < counts_[b]; ++i)
            bh.push(ids_[static_cast<std::size_t>(b) * r_ + i], static_cast<DistType>(dists_[static_cast<std::size_t>(b) * r_ + i]));
    }
};

#endif
