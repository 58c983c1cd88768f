// Minimal stand-in for the cereal serialisation library (absent from this image; the reference
// pins cereal 1.2.2 in getdeps.sh).  TEST INFRASTRUCTURE ONLY.
//
// It implements just enough of cereal's API for the reference's own save()/load() members
// (databases.hpp:158-167, :300-331; quantizers.hpp:170-187, :303-323) and load_database
// (query_common.hpp:321-328) to run: the ORDER of the fields in a file is therefore decided by
// unmodified reference code.  The BYTES of each field follow cereal 1.2.2's BinaryOutputArchive
// as published (native endianness, no header), restated here from its documentation/sources:
//   arithmetic value              raw sizeof(T) bytes
//   binary_data(p, n)             n raw bytes
//   std::vector<arithmetic>       uint64 element count, then the raw elements
//   std::string                   uint64 length, then the characters
//   std::unique_ptr<polymorphic>  uint32 polymorphic id:
//                                   0            null pointer
//                                   0x40000000   dynamic type == static type, no name follows
//                                   0x80000000|n first use of registered type n (n = 1, 2, .. in
//                                                order of first use in this archive), followed by
//                                                the registered name as a std::string
//                                   n            later use of registered type n
//                                 then the pointer wrapper: uint8 valid (1), then the object
//   base_class<B>(this)           B's fields, inline
// This restatement cannot be checked against the real library here; DESIGN.md says so.
#ifndef QADC_CEREAL_STUB_HPP
#define QADC_CEREAL_STUB_HPP
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <typeindex>
#include <typeinfo>
#include <utility>
#include <vector>

namespace cereal {

struct binary_blob { void* p; std::size_t n; };
template <class T> inline binary_blob binary_data(T* p, std::size_t n) {
    return binary_blob{const_cast<void*>(static_cast<const void*>(p)), n};
}
template <class B> struct base_ref { B* ptr; };
template <class B, class D> inline base_ref<B> base_class(D* d) {
    return base_ref<B>{const_cast<B*>(static_cast<const B*>(d))};
}

struct BinaryOutputArchive;
struct BinaryInputArchive;

// classes serialised through their own save()/load() members = every class that is not one of
// the wrappers/containers handled explicitly below
template <class T> struct is_wrapped : std::false_type {};
template <> struct is_wrapped<binary_blob> : std::true_type {};
template <> struct is_wrapped<std::string> : std::true_type {};
template <class B> struct is_wrapped<base_ref<B>> : std::true_type {};
template <class T> struct is_wrapped<std::vector<T>> : std::true_type {};
template <class T> struct is_wrapped<std::unique_ptr<T>> : std::true_type {};
template <class T> struct has_members : std::integral_constant<bool, std::is_class<T>::value && !is_wrapped<T>::value> {};

namespace mini {
const std::uint32_t kMsb = 0x80000000u, kMsb2 = 0x40000000u;
struct Entry {
    std::string name;
    std::function<void(BinaryOutputArchive&, const void*)> save;   // most-derived object
    std::function<void*(BinaryInputArchive&)> load;                // returns new most-derived object
};
inline std::map<std::type_index, Entry>& by_type() { static std::map<std::type_index, Entry> m; return m; }
inline std::map<std::string, std::type_index>& by_name() { static std::map<std::string, std::type_index> m; return m; }
typedef void* (*Upcast)(void*);
inline std::map<std::pair<std::type_index, std::type_index>, Upcast>& upcasts() {
    static std::map<std::pair<std::type_index, std::type_index>, Upcast> m;
    return m;
}
}  // namespace mini

struct BinaryOutputArchive {
    std::ostream& os;
    std::map<std::string, std::uint32_t> ids;
    std::uint32_t next_id = 1;
    explicit BinaryOutputArchive(std::ostream& s) : os(s) {}
    void raw(const void* p, std::size_t n) { os.write(static_cast<const char*>(p), static_cast<std::streamsize>(n)); }

    template <class... A> void operator()(A&&... a) {
        int order[] = {0, (put(a), 0)...};
        (void)order;
    }
    template <class T> typename std::enable_if<std::is_arithmetic<T>::value>::type put(const T& v) { raw(&v, sizeof(T)); }
    void put(const binary_blob& b) { raw(b.p, b.n); }
    void put(const std::string& s) {
        const std::uint64_t n = s.size();
        raw(&n, 8);
        raw(s.data(), s.size());
    }
    template <class T> typename std::enable_if<std::is_arithmetic<T>::value>::type put(const std::vector<T>& v) {
        const std::uint64_t n = v.size();
        raw(&n, 8);
        raw(v.data(), v.size() * sizeof(T));
    }
    template <class B> void put(const base_ref<B>& b) { static_cast<const B*>(b.ptr)->save(*this); }
    template <class T> typename std::enable_if<has_members<T>::value>::type put(const T& obj) { obj.save(*this); }
    template <class T> void put(const std::unique_ptr<T>& ptr) {
        static_assert(std::is_polymorphic<T>::value, "only polymorphic unique_ptr is used by the reference");
        if (!ptr) { put(std::uint32_t(0)); return; }
        if (typeid(*ptr) == typeid(T)) {
            put(mini::kMsb2);
            put(std::uint8_t(1));
            save_same(*ptr, std::integral_constant<bool, std::is_abstract<T>::value>());
            return;
        }
        auto it = mini::by_type().find(std::type_index(typeid(*ptr)));
        if (it == mini::by_type().end()) throw std::runtime_error("cereal-mini: unregistered polymorphic type");
        const mini::Entry& e = it->second;
        auto id = ids.find(e.name);
        if (id == ids.end()) {
            const std::uint32_t n = next_id++;
            ids[e.name] = n;
            put(n | mini::kMsb);
            put(e.name);
        } else {
            put(id->second);
        }
        put(std::uint8_t(1));
        e.save(*this, dynamic_cast<const void*>(ptr.get()));
    }
    template <class T> void save_same(const T& obj, std::false_type) { obj.save(*this); }
    template <class T> void save_same(const T&, std::true_type) {}
};

struct BinaryInputArchive {
    std::istream& is;
    std::map<std::uint32_t, std::string> names;
    explicit BinaryInputArchive(std::istream& s) : is(s) {}
    void raw(void* p, std::size_t n) {
        is.read(static_cast<char*>(p), static_cast<std::streamsize>(n));
        if (!is) throw std::runtime_error("cereal-mini: truncated archive");
    }

    template <class... A> void operator()(A&&... a) {
        int order[] = {0, (get(a), 0)...};
        (void)order;
    }
    template <class T> typename std::enable_if<std::is_arithmetic<T>::value>::type get(T& v) { raw(&v, sizeof(T)); }
    void get(const binary_blob& b) { raw(b.p, b.n); }
    void get(std::string& s) {
        std::uint64_t n;
        raw(&n, 8);
        s.resize(n);
        if (n) raw(&s[0], n);
    }
    template <class T> typename std::enable_if<std::is_arithmetic<T>::value>::type get(std::vector<T>& v) {
        std::uint64_t n;
        raw(&n, 8);
        v.resize(n);
        raw(v.data(), n * sizeof(T));
    }
    template <class B> void get(const base_ref<B>& b) { b.ptr->load(*this); }
    template <class T> typename std::enable_if<has_members<T>::value>::type get(T& obj) { obj.load(*this); }
    template <class T> void get(std::unique_ptr<T>& ptr) {
        std::uint32_t id;
        get(id);
        if (id == 0) { ptr.reset(); return; }
        if (id & mini::kMsb2) {
            ptr.reset(load_same<T>(std::integral_constant<bool, std::is_abstract<T>::value>()));
            return;
        }
        std::string name;
        if (id & mini::kMsb) {
            get(name);
            names[id & ~mini::kMsb] = name;
        } else {
            auto it = names.find(id);
            if (it == names.end()) throw std::runtime_error("cereal-mini: unknown polymorphic id");
            name = it->second;
        }
        auto t = mini::by_name().find(name);
        if (t == mini::by_name().end()) throw std::runtime_error("cereal-mini: unregistered type name " + name);
        auto up = mini::upcasts().find(std::make_pair(std::type_index(typeid(T)), t->second));
        if (up == mini::upcasts().end()) throw std::runtime_error("cereal-mini: no relation to " + name);
        ptr.reset(static_cast<T*>(up->second(mini::by_type().at(t->second).load(*this))));
    }
    template <class T> T* load_same(std::false_type) {
        std::uint8_t valid;
        get(valid);
        if (!valid) return nullptr;
        std::unique_ptr<T> obj(new T());
        obj->load(*this);
        return obj.release();
    }
    template <class T> T* load_same(std::true_type) { throw std::runtime_error("cereal-mini: abstract type in archive"); }
};

namespace mini {
template <class T> struct Registrar {
    explicit Registrar(const char* name) {
        Entry e;
        e.name = name;
        e.save = [](BinaryOutputArchive& ar, const void* p) { static_cast<const T*>(p)->save(ar); };
        e.load = [](BinaryInputArchive& ar) -> void* {
            std::uint8_t valid;
            ar.get(valid);
            if (!valid) return nullptr;
            std::unique_ptr<T> obj(new T());
            obj->load(ar);
            return obj.release();
        };
        by_type().emplace(std::type_index(typeid(T)), e);
        by_name().emplace(std::string(name), std::type_index(typeid(T)));
    }
};
template <class B, class D> struct Relation {
    Relation() {
        upcasts()[std::make_pair(std::type_index(typeid(B)), std::type_index(typeid(D)))] =
            [](void* d) -> void* { return static_cast<B*>(static_cast<D*>(d)); };
    }
};
}  // namespace mini
}  // namespace cereal

#define CEREAL_REGISTER_TYPE(T) static ::cereal::mini::Registrar<T> qadc_cereal_reg_##T(#T)
#define CEREAL_REGISTER_POLYMORPHIC_RELATION(B, D) static ::cereal::mini::Relation<B, D> qadc_cereal_rel_##B##_##D
#endif
