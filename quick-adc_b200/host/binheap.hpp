// kv_binheap<Key,Value> — bounded max-heap holding the k smallest values, same public
// surface as the reference's binheap.hpp:18-142 (push, max, keys, values, size, capacity,
// sort, reset, reset_capacity) so that process_queries and the recall check read unchanged.
// Own implementation on std::push_heap-free sift loops; semantics kept: not full -> append and
// sift up (strict >); full -> replace the root iff value < root (strict).
#ifndef QADC_HOST_BINHEAP_HPP_
#define QADC_HOST_BINHEAP_HPP_

#include <algorithm>
#include <numeric>
#include <utility>
#include <vector>

template <typename KeyType, typename ValueType>
class kv_binheap {
    std::vector<KeyType> keys_;
    std::vector<ValueType> values_;
    int capacity_ = 0;
    int size_ = 0;

    void swap_nodes(int a, int b) {
        std::swap(values_[a], values_[b]);
        std::swap(keys_[a], keys_[b]);
    }

public:
    kv_binheap() = default;
    explicit kv_binheap(int capacity) { reset_capacity(capacity); }

    void reset_capacity(int capacity) {
        capacity_ = capacity;
        size_ = 0;
        keys_.assign(capacity, KeyType());
        values_.assign(capacity, ValueType());
    }
    void reset() { size_ = 0; }
    int capacity() const { return capacity_; }
    int size() const { return size_; }
    ValueType max() const { return values_[0]; }
    const KeyType* keys() const { return keys_.data(); }
    const ValueType* values() const { return values_.data(); }

    void push(KeyType key, ValueType value) {
        if (size_ != capacity_) {
            int node = size_++;
            values_[node] = value;
            keys_[node] = key;
            while (node != 0) {
                const int parent = (node - 1) / 2;
                if (!(values_[node] > values_[parent])) break;
                swap_nodes(node, parent);
                node = parent;
            }
            return;
        }
        if (!(value < values_[0])) return;
        values_[0] = value;
        keys_[0] = key;
        int node = 0;
        for (;;) {
            const int left = 2 * node + 1, right = left + 1;
            if (left >= size_) break;
            int child = left;
            if (right < size_ && values_[right] > values_[left]) child = right;
            if (values_[child] <= values_[node]) break;
            swap_nodes(node, child);
            node = child;
        }
    }

    // ascending by value (like the reference: non-stable, value only)
    void sort(KeyType keys[], ValueType values[]) const {
        std::vector<int> order(size_);
        std::iota(order.begin(), order.end(), 0);
        std::sort(order.begin(), order.end(), [this](int a, int b) { return values_[a] < values_[b]; });
        for (int i = 0; i < size_; ++i) {
            keys[i] = keys_[order[i]];
            values[i] = values_[order[i]];
        }
    }
};

#endif
