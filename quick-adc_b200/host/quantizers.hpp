// base_pq / opq — the quantiser state the reference keeps (quantizers.hpp:96-168, :248-301) and
// its .pq.data / .opq.data reader (quantizers.cpp:27-46: int32 dim, m, bits; float codebooks
// [dim * 2^bits]; opq: float rotation[dim*dim]).  Encoding (add_vectors) is a "next" row.
#ifndef QADC_HOST_QUANTIZERS_HPP_
#define QADC_HOST_QUANTIZERS_HPP_

#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

struct base_pq {
    int sq_count = 0;
    int sq_bits = 0;
    int dim = 0;
    std::vector<float> centroids_flat;   // sq_count x 2^bits x sq_dim

    base_pq() = default;
    base_pq(int sq_count_, int sq_bits_, int dim_, const float* centroids = nullptr)
        : sq_count(sq_count_), sq_bits(sq_bits_), dim(dim_), centroids_flat(static_cast<size_t>(dim_) << sq_bits_) {
        if (centroids) std::copy(centroids, centroids + centroids_flat.size(), centroids_flat.begin());
    }
    virtual ~base_pq() = default;
    virtual std::string get_tag() const { return "pq"; }
    virtual const float* rotation_ptr() const { return nullptr; }
    int sq_dim() const { return dim / sq_count; }
    int sq_centroid_count() const { return 1 << sq_bits; }
    int all_centroids_dim() const { return dim * sq_centroid_count(); }
    int code_size() const { return sq_count * sq_bits / 8; }
    const float* centroids(int sq_i) const { return centroids_flat.data() + static_cast<size_t>(sq_i) * sq_centroid_count() * sq_dim(); }
    void print(std::ostream& os) const { os << get_tag() << " (dim=" << dim << ", sq=" << sq_count << "x" << sq_bits << ")"; }
};

struct opq : base_pq {
    std::vector<float> rotation;   // dim x dim, row-major; vectors are rotated as X * R^T
    opq() = default;
    opq(int sq_count_, int sq_bits_, int dim_, const float* centroids = nullptr, const float* rotation_ = nullptr)
        : base_pq(sq_count_, sq_bits_, dim_, centroids), rotation(static_cast<size_t>(dim_) * dim_) {
        if (rotation_) std::copy(rotation_, rotation_ + rotation.size(), rotation.begin());
    }
    std::string get_tag() const override { return "opq"; }
    const float* rotation_ptr() const override { return rotation.data(); }
};

inline bool qadc_ends_with(const std::string& s, const std::string& e) {
    return s.size() >= e.size() && std::equal(e.rbegin(), e.rend(), s.rbegin());
}

// quantizers.cpp:27-46 / README.md:344-362
inline std::unique_ptr<base_pq> pq_from_data_file(const char* filename) {
    std::ifstream in(filename, std::ios::binary);
    if (!in) {
        std::cerr << "Could not open " << filename << std::endl;
        std::exit(1);
    }
    std::int32_t hdr[3];
    in.read(reinterpret_cast<char*>(hdr), sizeof(hdr));
    const int dim = hdr[0], m = hdr[1], bits = hdr[2];
    const std::string name(filename);
    std::unique_ptr<base_pq> pq;
    if (qadc_ends_with(name, ".opq.data")) pq.reset(new opq(m, bits, dim));
    else pq.reset(new base_pq(m, bits, dim));
    in.read(reinterpret_cast<char*>(pq->centroids_flat.data()), pq->centroids_flat.size() * sizeof(float));
    if (auto* o = dynamic_cast<opq*>(pq.get()))
        in.read(reinterpret_cast<char*>(o->rotation.data()), o->rotation.size() * sizeof(float));
    if (!in) {
        std::cerr << "Truncated quantizer file " << filename << std::endl;
        std::exit(1);
    }
    return pq;
}

#endif
