// base_db / flat_db / index_db — the database API the scanner consumes (databases.hpp:34-63,
// :77-134, :176-250): partition_count / get_partition / free_partition, row-major 4-bit codes,
// uint32 labels for inverted lists.  Query-time assignment and residuals run on the GPU, so
// assign_compute_residuals* are not needed here.
//
// Files.  load_database() reads two formats, told apart by their first bytes:
//  * the reference's own database files (cereal 1.2.2 BinaryOutputArchive of a
//    std::unique_ptr<base_db>, flatdb_create.cpp:49-53 / query_common.hpp:321-328) — layout in
//    the "cereal archive" section below;
//  * the ".qdb" container (writer: quick-adc_b200/dbfile.py), a plain header + arrays:
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
    virtual void save(std::ostream& os) const = 0;          // .qdb container
    virtual void save_archive(std::ostream& os) const = 0;  // the reference's cereal layout
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
    void save_archive(std::ostream& os) const override;
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
    void save_archive(std::ostream& os) const override;
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

inline std::unique_ptr<base_db> load_qdb(std::istream& in, const char* filename, std::uint64_t file_bytes) {
    char magic[8];
    std::int32_t h[6];
    in.read(magic, 8);
    in.read(reinterpret_cast<char*>(h), sizeof(h));
    if (!in || std::memcmp(magic, "QADCDB1", 8) != 0) {
        std::cerr << filename << " is not a .qdb database" << std::endl;
        std::exit(1);
    }
    const int kind = h[0], pq_kind = h[1], dim = h[2], m = h[3], bits = h[4], K = h[5];
    // every header field is checked before it sizes an allocation (a corrupt file must say so, not crash):
    // geometry like get_pq below, every array bounded by the file length
    const bool geometry_ok = (kind == 0 || kind == 1) && (pq_kind == 0 || pq_kind == 1) && m > 0 && (bits == 4 || bits == 8 || bits == 16) &&
                             dim > 0 && dim <= (1 << 20) && dim % m == 0 && (m * bits) % 8 == 0 && K > 0 && (kind == 1 || K == 1);
    const std::uint64_t fixed_bytes = geometry_ok ? ((static_cast<std::uint64_t>(dim) << bits) + (pq_kind ? static_cast<std::uint64_t>(dim) * dim : 0) +
                                                     (kind ? static_cast<std::uint64_t>(K) * dim : 0)) * 4 + static_cast<std::uint64_t>(K) * 8
                                                  : 0;
    if (!geometry_ok || fixed_bytes > file_bytes) {
        std::cerr << filename << ": corrupt .qdb header" << std::endl;
        std::exit(1);
    }
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
        std::uint64_t size = 0;
        in.read(reinterpret_cast<char*>(&size), 8);
        if (!in || size > file_bytes / cs || size >= (std::uint64_t(1) << 32)) {
            std::cerr << filename << ": corrupt .qdb size field" << std::endl;
            std::exit(1);
        }
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
        std::uint64_t total = 0;
        for (int p = 0; p < K && in; ++p) total = (sizes[p] > file_bytes) ? file_bytes + 1 : total + sizes[p];
        if (!in || total > file_bytes / (cs + 4)) {
            std::cerr << filename << ": corrupt .qdb partition sizes" << std::endl;
            std::exit(1);
        }
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

// ---- cereal archive -------------------------------------------------------------------------
// The reference writes `std::unique_ptr<base_db>` through cereal::BinaryOutputArchive (cereal
// 1.2.2: native little-endian values, no header).  Bytes, in order:
//   uint32 0x80000001; uint64 len; "flat_db" | "index_db"        first registered polymorphic type
//   uint8  1                                                      pointer is not null
//   flat_db  (databases.hpp:158-161):  <pq>; uint32 codes_count; uint64 n; uint8 codes[n]
//   index_db (databases.hpp:300-313):  int32 part_count; <pq>; float centroids[part_count*dim];
//                                      per partition uint64 n; uint8 codes[n];
//                                      per partition uint64 n; uint32 labels[n]
//   <pq> = std::unique_ptr<base_pq>:
//     base_pq: uint32 0x40000000; uint8 1; int32 sq_count, sq_bits, dim; float codebooks[dim*2^bits]
//     opq:     uint32 0x80000002; uint64 3; "opq"; uint8 1; the base_pq fields; float rotation[dim*dim]
// (quantizers.hpp:170-178, :303-312).  cereal is not available in this build environment: the
// layout is restated from its published format and checked against the reference's own
// save()/load() code running over a restatement of the archive classes (oracle/shims/cereal),
// not against files written by the real library.
namespace qadc_archive {
const std::uint32_t kFirstUse = 0x80000000u, kSameType = 0x40000000u;

template <typename T> inline void put(std::ostream& os, const T& v) { os.write(reinterpret_cast<const char*>(&v), sizeof(T)); }
template <typename T> inline T get(std::istream& in) {
    T v = T();
    in.read(reinterpret_cast<char*>(&v), sizeof(T));
    return v;
}
inline void put_name(std::ostream& os, std::uint32_t id, const char* name) {
    put<std::uint32_t>(os, kFirstUse | id);
    put<std::uint64_t>(os, std::strlen(name));
    os.write(name, static_cast<std::streamsize>(std::strlen(name)));
    put<std::uint8_t>(os, 1);
}
template <typename T> inline void put_vector(std::ostream& os, const std::vector<T>& v) {
    put<std::uint64_t>(os, v.size());
    os.write(reinterpret_cast<const char*>(v.data()), static_cast<std::streamsize>(v.size() * sizeof(T)));
}
template <typename T> inline bool get_vector(std::istream& in, std::vector<T>& v, std::uint64_t max_bytes) {
    const std::uint64_t n = get<std::uint64_t>(in);
    if (!in || n > max_bytes / sizeof(T)) return false;
    v.resize(n);
    in.read(reinterpret_cast<char*>(v.data()), static_cast<std::streamsize>(n * sizeof(T)));
    return static_cast<bool>(in);
}
inline std::string get_name(std::istream& in) {
    const std::uint64_t n = get<std::uint64_t>(in);
    if (!in || n > 64) return std::string();
    std::string s(n, '\0');
    in.read(&s[0], static_cast<std::streamsize>(n));
    return s;
}
inline void put_pq(std::ostream& os, const base_pq& pq, std::uint32_t next_id) {
    const opq* o = dynamic_cast<const opq*>(&pq);
    if (o) put_name(os, next_id, "opq");
    else { put<std::uint32_t>(os, kSameType); put<std::uint8_t>(os, 1); }
    put<std::int32_t>(os, pq.sq_count);
    put<std::int32_t>(os, pq.sq_bits);
    put<std::int32_t>(os, pq.dim);
    os.write(reinterpret_cast<const char*>(pq.centroids_flat.data()), static_cast<std::streamsize>(pq.centroids_flat.size() * sizeof(float)));
    if (o) os.write(reinterpret_cast<const char*>(o->rotation.data()), static_cast<std::streamsize>(o->rotation.size() * sizeof(float)));
}
inline std::unique_ptr<base_pq> get_pq(std::istream& in) {
    const std::uint32_t id = get<std::uint32_t>(in);
    bool is_opq = false;
    if (id & kFirstUse) {
        if (get_name(in) != "opq") return nullptr;
        is_opq = true;
    } else if (!(id & kSameType)) {
        return nullptr;
    }
    if (get<std::uint8_t>(in) != 1) return nullptr;
    const int m = get<std::int32_t>(in), bits = get<std::int32_t>(in), dim = get<std::int32_t>(in);
    if (!in || m <= 0 || bits <= 0 || bits > 16 || dim <= 0 || dim > (1 << 20) || dim % m) return nullptr;
    std::unique_ptr<base_pq> pq;
    if (is_opq) pq.reset(new opq(m, bits, dim));
    else pq.reset(new base_pq(m, bits, dim));
    in.read(reinterpret_cast<char*>(pq->centroids_flat.data()), static_cast<std::streamsize>(pq->centroids_flat.size() * sizeof(float)));
    if (auto* o = dynamic_cast<opq*>(pq.get()))
        in.read(reinterpret_cast<char*>(o->rotation.data()), static_cast<std::streamsize>(o->rotation.size() * sizeof(float)));
    if (!in) return nullptr;
    return pq;
}
}  // namespace qadc_archive

inline void flat_db::save_archive(std::ostream& os) const {
    qadc_archive::put_name(os, 1, "flat_db");
    qadc_archive::put_pq(os, *pq, 2);
    qadc_archive::put<std::uint32_t>(os, codes_count);
    qadc_archive::put_vector(os, codes);
}

inline void index_db::save_archive(std::ostream& os) const {
    qadc_archive::put_name(os, 1, "index_db");
    qadc_archive::put<std::int32_t>(os, part_count);
    qadc_archive::put_pq(os, *pq, 2);
    os.write(reinterpret_cast<const char*>(centroids.data()), static_cast<std::streamsize>(centroids.size() * sizeof(float)));
    for (int p = 0; p < part_count; ++p) qadc_archive::put_vector(os, partitions[p]);
    for (int p = 0; p < part_count; ++p) qadc_archive::put_vector(os, labels[p]);
}

// `file_bytes` bounds every length field, so a corrupt file cannot ask for absurd allocations.
inline std::unique_ptr<base_db> load_archive(std::istream& in, std::uint64_t file_bytes) {
    using namespace qadc_archive;
    const std::uint32_t id = get<std::uint32_t>(in);
    if (!in || !(id & kFirstUse)) return nullptr;
    const std::string name = get_name(in);
    if (get<std::uint8_t>(in) != 1) return nullptr;
    std::unique_ptr<base_db> db;
    if (name == "flat_db") {
        auto* f = new flat_db;
        db.reset(f);
        f->pq = get_pq(in);
        if (!f->pq) return nullptr;
        f->codes_count = get<std::uint32_t>(in);
        if (!get_vector(in, f->codes, file_bytes)) return nullptr;
        if (f->codes.size() != static_cast<size_t>(f->codes_count) * f->pq->code_size()) return nullptr;
    } else if (name == "index_db") {
        auto* x = new index_db;
        db.reset(x);
        x->part_count = get<std::int32_t>(in);
        x->pq = get_pq(in);
        if (!x->pq || x->part_count <= 0 || static_cast<std::uint64_t>(x->part_count) * x->pq->dim * 4 > file_bytes) return nullptr;
        x->centroids.resize(static_cast<size_t>(x->part_count) * x->pq->dim);
        in.read(reinterpret_cast<char*>(x->centroids.data()), static_cast<std::streamsize>(x->centroids.size() * sizeof(float)));
        x->partitions.resize(x->part_count);
        x->labels.resize(x->part_count);
        for (int p = 0; p < x->part_count; ++p)
            if (!get_vector(in, x->partitions[p], file_bytes)) return nullptr;
        for (int p = 0; p < x->part_count; ++p) {
            if (!get_vector(in, x->labels[p], file_bytes)) return nullptr;
            if (x->partitions[p].size() != x->labels[p].size() * x->pq->code_size()) return nullptr;
        }
    } else {
        return nullptr;
    }
    return db;
}

// query_common.hpp:321-328 — accepts the reference's archives and .qdb containers.
inline std::unique_ptr<base_db> load_database(const char* filename) {
    std::ifstream in(filename, std::ios::binary);
    if (!in) {
        std::cerr << "Could not open database " << filename << std::endl;
        std::exit(1);
    }
    in.seekg(0, std::ios::end);
    const std::uint64_t file_bytes = static_cast<std::uint64_t>(in.tellg());
    in.seekg(0);
    char magic[8] = {0};
    in.read(magic, 8);
    in.clear();
    in.seekg(0);
    if (std::memcmp(magic, "QADCDB1", 8) == 0) return load_qdb(in, filename, file_bytes);
    std::unique_ptr<base_db> db = load_archive(in, file_bytes);
    if (!db) {
        std::cerr << filename << " is not a database file (neither a flat_db/index_db cereal archive nor a .qdb container)" << std::endl;
        std::exit(1);
    }
    return db;
}

// Writes `.qdb` when the name ends so, the reference's archive layout otherwise.
inline bool save_database(const base_db& db, const char* filename) {
    std::ofstream out(filename, std::ios::binary);
    if (!out) return false;
    if (qadc_ends_with(filename, ".qdb")) db.save(out);
    else db.save_archive(out);
    return static_cast<bool>(out);
}

#endif
