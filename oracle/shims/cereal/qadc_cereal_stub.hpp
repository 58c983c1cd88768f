// No-op stand-in for the cereal serialisation library (absent from this image; the
// reference pins cereal 1.2.2 in getdeps.sh). The oracle builds databases in memory,
// so archives only need to compile, never to run. TEST INFRASTRUCTURE ONLY.
#ifndef QADC_CEREAL_STUB_HPP
#define QADC_CEREAL_STUB_HPP
#include <cstddef>
#include <cstdint>
#include <algorithm>
#include <iostream>
#include <fstream>
#include <memory>
#include <string>
#include <vector>
namespace cereal {
struct binary_blob { void* p; std::size_t n; };
template <class T> inline binary_blob binary_data(T* p, std::size_t n) {
    return binary_blob{const_cast<void*>(static_cast<const void*>(p)), n};
}
template <class B> struct base_ref { const void* d; };
template <class B, class D> inline base_ref<B> base_class(D* d) { return base_ref<B>{d}; }
struct BinaryInputArchive {
    explicit BinaryInputArchive(std::istream&) {}
    template <class... A> void operator()(A&&...) {}
};
struct BinaryOutputArchive {
    explicit BinaryOutputArchive(std::ostream&) {}
    template <class... A> void operator()(A&&...) {}
};
}  // namespace cereal
#define CEREAL_REGISTER_TYPE(T)
#define CEREAL_REGISTER_POLYMORPHIC_RELATION(B, D)
#endif
