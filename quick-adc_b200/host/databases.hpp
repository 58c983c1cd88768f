// base_db / flat_db / index_db — the database API the scanner consumes (databases.hpp:34-63,
// :77-134, :176-250): partition_count / get_partition / free_partition, row-major 4-bit codes,
// uint32 labels for inverted lists.  Query-time assignment and residuals run on the GPU, so
// assign_compute_residuals* are not needed here.  Files: the reference serialises with cereal
// (absent here, layout unverifiable), so databases are stored in the small documented ".qdb"
// container below (writer: quick-adc_b200/dbfile.py).
//
//   char  magic[8] = "QADCDB1\0"
//   int32 kind (0 flat, 1 index), pq_kind (0 pq, 1 opq), dim, m, bits, K (partitions; 1 if flat)
//   float codebooks[dim * 2^bits]; [opq] float rotation[dim*dim]; [index] float centroids[K*dim]
//   uint64 sizes[K]
//   per partition: uint8 codes[size * m*bits/8]; [index] uint32 labels[size]
#ifndef QADC_HOST_DATABASES_HPP_
#define QADC_HOST_DATABASES_HPP_

#include <cstring>

#include "../../include/qadc_b200.h"
#include "quantizers.hpp"

struct base_db {
    std::unique_ptr<base_pq> pq;
    virtual ~base_db() = default;
    // base_db::add_vectors (databases.hpp:57-58): encode `count` vectors and store them with ids
    // labels_offset.. ; the PQ encoding (and, for inverted lists, the coarse assignment) runs on
    // the GPU through `enc`, a context prepared with qadc_set_pq (+ qadc_set_coarse).
    virtual void add_vectors(qadc_ctx* enc, const float* vectors, unsigned count, unsigned labels_offset) = 0;
    virtual void save(std::ostream& os) const = 0;
    virtual int partition_count() const = 0;
    virtual void get_partition(int part_i, const std::uint8_t*& codes, unsigned*& labels, unsigned& size) const = 0;
    virtual void free_partition(int part_i) = 0;
    virtual const float* coarse_centroids() const { return nullptr; }   // index_db::centroids
    virtual void print(std::ostream& os) const = 0;
};

struct flat_db : base_db {
    std::vector<std::uint8_t> codes;
    unsigned codes_count = 0;
    int partition_count() const override { return 1; }
    void get_partition(int, const std::uint8_t*& codes_, unsigned*& labels, unsigned& size) const override {
        codes_ = codes.data();
        labels = nullptr;
        size = codes_count;
    }
    void free_partition(int) override {
        std::vector<std::uint8_t>().swap(codes);
        codes_count = 0;
    }
    void print(std::ostream& os) const override { os << "Flat DB" << std::endl; pq->print(os); }
    // flat_db::add_vectors (databases.hpp:136-156): codes are stored at positions labels_offset..
    void add_vectors(qadc_ctx* enc, const float* vectors, unsigned count, unsigned labels_offset) override {
        const size_t cs = pq->code_size();
        if (labels_offset + count > codes_count) {
            codes_count = labels_offset + count;
            codes.resize(static_cast<size_t>(codes_count) * cs);
        }
        if (qadc_encode(enc, vectors, count, nullptr, codes.data() + static_cast<size_t>(labels_offset) * cs)) {
            std::cerr << "qadc_encode: " << qadc_last_error(enc) << std::endl;
            std::exit(1);
        }
    }
    void save(std::ostream& os) const override;
};

struct index_db : base_db {
    int part_count = 0;
    std::vector<float> centroids;   // part_count x dim
    std::vector<std::vector<std::uint8_t>> partitions;
    std::vector<std::vector<unsigned>> labels;
    int partition_count() const override { return part_count; }
    void get_partition(int p, const std::uint8_t*& codes_, unsigned*& labels_, unsigned& size) const override {
        codes_ = partitions[p].data();
        labels_ = const_cast<unsigned*>(labels[p].data());
        size = static_cast<unsigned>(labels[p].size());
    }
    void free_partition(int p) override {
        std::vector<std::uint8_t>().swap(partitions[p]);
        std::vector<unsigned>().swap(labels[p]);
    }
    const float* coarse_centroids() const override { return centroids.data(); }
    void print(std::ostream& os) const override {
        os << "Indexed DB (partitions=" << part_count << ")" << std::endl;
        pq->print(os);
    }
    // index_db::add_vectors (databases.hpp:270-298): nearest cell, code of the residual, dispatch
    // into the cell's list with label vec_i + labels_offset.
    void add_vectors(qadc_ctx* enc, const float* vectors, unsigned count, unsigned labels_offset) override {
        const size_t cs = pq->code_size();
        std::vector<std::int32_t> assign(count);
        std::vector<std::uint8_t> buf(static_cast<size_t>(count) * cs);
        if (qadc_encode(enc, vectors, count, assign.data(), buf.data())) {
            std::cerr << "qadc_encode: " << qadc_last_error(enc) << std::endl;
            std::exit(1);
        }
        for (unsigned v = 0; v < count; ++v) {
            const int p = assign[v];
            partitions[p].insert(partitions[p].end(), buf.begin() + v * cs, buf.begin() + (v + 1) * cs);
            labels[p].push_back(v + labels_offset);
        }
    }
    void save(std::ostream& os) const override;
};

inline void qdb_write_header(std::ostream& os, const base_pq& pq, int kind, int K) {
    const opq* o = dynamic_cast<const opq*>(&pq);
    const std::int32_t h[6] = {kind, o ? 1 : 0, pq.dim, pq.sq_count, pq.sq_bits, K};
    os.write("QADCDB1\0", 8);
    os.write(reinterpret_cast<const char*>(h), sizeof(h));
    os.write(reinterpret_cast<const char*>(pq.centroids_flat.data()), pq.centroids_flat.size() * sizeof(float));
    if (o) os.write(reinterpret_cast<const char*>(o->rotation.data()), o->rotation.size() * sizeof(float));
}

inline void flat_db::save(std::ostream& os) const {
    qdb_write_header(os, *pq, 0, 1);
    const std::uint64_t n = codes_count;
    os.write(reinterpret_cast<const char*>(&n), 8);
    os.write(reinterpret_cast<const char*>(codes.data()), codes.size());
}

inline void index_db::save(std::ostream& os) const {
    qdb_write_header(os, *pq, 1, part_count);
    os.write(reinterpret_cast<const char*>(centroids.data()), centroids.size() * sizeof(float));
    for (int p = 0; p < part_count; ++p) {
        const std::uint64_t n = labels[p].size();
        os.write(reinterpret_cast<const char*>(&n), 8);
    }
    for (int p = 0; p < part_count; ++p) {
        os.write(reinterpret_cast<const char*>(partitions[p].data()), partitions[p].size());
        os.write(reinterpret_cast<const char*>(labels[p].data()), labels[p].size() * 4);
    }
}

inline std::unique_ptr<base_db> load_qdb(const char* filename) {
    std::ifstream in(filename, std::ios::binary);
    if (!in) {
        std::cerr << "Could not open database " << filename << std::endl;
        std::exit(1);
    }
    char magic[8];
    std::int32_t h[6];
    in.read(magic, 8);
    in.read(reinterpret_cast<char*>(h), sizeof(h));
    if (!in || std::memcmp(magic, "QADCDB1", 8) != 0) {
        std::cerr << filename << " is not a .qdb database (cereal archives of the reference are not readable "
                  << "here: convert with quick-adc_b200/dbfile.py)" << std::endl;
        std::exit(1);
    }
    const int kind = h[0], pq_kind = h[1], dim = h[2], m = h[3], bits = h[4], K = h[5];
    std::unique_ptr<base_pq> pq;
    if (pq_kind == 1) pq.reset(new opq(m, bits, dim));
    else pq.reset(new base_pq(m, bits, dim));
    in.read(reinterpret_cast<char*>(pq->centroids_flat.data()), pq->centroids_flat.size() * sizeof(float));
    if (auto* o = dynamic_cast<opq*>(pq.get())) in.read(reinterpret_cast<char*>(o->rotation.data()), o->rotation.size() * sizeof(float));
    const size_t cs = static_cast<size_t>(m) * bits / 8;
    std::unique_ptr<base_db> db;
    if (kind == 0) {
        auto* f = new flat_db;
        db.reset(f);
        std::uint64_t size;
        in.read(reinterpret_cast<char*>(&size), 8);
        f->codes.resize(size * cs);
        in.read(reinterpret_cast<char*>(f->codes.data()), f->codes.size());
        f->codes_count = static_cast<unsigned>(size);
    } else {
        auto* x = new index_db;
        db.reset(x);
        x->part_count = K;
        x->centroids.resize(static_cast<size_t>(K) * dim);
        in.read(reinterpret_cast<char*>(x->centroids.data()), x->centroids.size() * sizeof(float));
        std::vector<std::uint64_t> sizes(K);
        in.read(reinterpret_cast<char*>(sizes.data()), 8 * static_cast<size_t>(K));
        x->partitions.resize(K);
        x->labels.resize(K);
        for (int p = 0; p < K; ++p) {
            x->partitions[p].resize(sizes[p] * cs);
            x->labels[p].resize(sizes[p]);
            in.read(reinterpret_cast<char*>(x->partitions[p].data()), x->partitions[p].size());
            in.read(reinterpret_cast<char*>(x->labels[p].data()), sizes[p] * 4);
        }
    }
    if (!in) {
        std::cerr << "Truncated database file " << filename << std::endl;
        std::exit(1);
    }
    db->pq = std::move(pq);
    return db;
}

#endif
