// Minimal .fvecs / .ivecs / .bvecs loaders (vector_io.cpp:40-58 behaviour: every vector is
// int32 dimension + payload; .bvecs bytes are widened to float).  The streaming reader the
// reference uses for db_add is out of scope.
#ifndef QADC_HOST_VECTOR_IO_HPP_
#define QADC_HOST_VECTOR_IO_HPP_

#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

template <typename T>
struct vectors_owner {
    std::unique_ptr<T[]> data;
    int dimension = 0;
    long count = 0;
    T* get(long vector_i) const { return data.get() + vector_i * dimension; }
};

template <typename T, typename FileT>
vectors_owner<T> load_vecs_as(const char* filename) {
    std::ifstream in(filename, std::ios::binary | std::ios::ate);
    if (!in) {
        std::cerr << "Could not open " << filename << std::endl;
        std::exit(1);
    }
    const long bytes = in.tellg();
    in.seekg(0);
    std::int32_t dim = 0;
    in.read(reinterpret_cast<char*>(&dim), 4);
    const long rec = 4 + static_cast<long>(dim) * sizeof(FileT);
    if (dim <= 0 || bytes % rec != 0) {
        std::cerr << "Malformed vecs file " << filename << std::endl;
        std::exit(1);
    }
    vectors_owner<T> v;
    v.dimension = dim;
    v.count = bytes / rec;
    v.data.reset(new T[v.count * dim]);
    std::vector<FileT> buf(dim);
    in.seekg(0);
    for (long i = 0; i < v.count; ++i) {
        std::int32_t d;
        in.read(reinterpret_cast<char*>(&d), 4);
        in.read(reinterpret_cast<char*>(buf.data()), dim * sizeof(FileT));
        for (int j = 0; j < dim; ++j) v.get(i)[j] = static_cast<T>(buf[j]);
    }
    return v;
}

inline bool vecs_has_ext(const std::string& s, const char* e) {
    const std::string x(e);
    return s.size() >= x.size() && s.compare(s.size() - x.size(), x.size(), x) == 0;
}

inline vectors_owner<float> load_vectors_by_extension(const char* filename) {
    const std::string s(filename);
    if (vecs_has_ext(s, ".fvecs")) return load_vecs_as<float, float>(filename);
    if (vecs_has_ext(s, ".bvecs")) return load_vecs_as<float, std::uint8_t>(filename);
    if (vecs_has_ext(s, ".ivecs")) return load_vecs_as<float, std::int32_t>(filename);
    std::cerr << "Unknown vector file extension: " << filename << std::endl;
    std::exit(1);
}

#endif
