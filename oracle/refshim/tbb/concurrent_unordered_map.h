// Stand-in for <tbb/concurrent_unordered_map.h> (see README.md).
//
// concurrent_unordered_multimap: all entries of a key are adjacent in iteration, NEWEST FIRST (TBB <= 2020
// internal_insert links a new node in front of the first equal one).  Distinct keys come in first-insertion
// order here (TBB: split order of the hash); no consumer depends on that order.
#pragma once
#include <atomic>
#include <cstddef>
#include <deque>
#include <unordered_map>
#include <utility>
#include <vector>
namespace tbb {

template <class T> class atomic {
public:
    atomic() : v_(T()) {}
    atomic(T v) : v_(v) {}
    atomic(const atomic &o) : v_(o.v_) {}
    atomic &operator=(T v) { v_ = v; return *this; }
    operator T() const { return v_; }
    T operator++(int) { return v_++; }
    T operator++() { return ++v_; }
private:
    T v_;
};

template <class K, class V, class H = std::hash<K>> class concurrent_unordered_map : public std::unordered_map<K, V, H> {
public:
    struct range_type {
        typename std::unordered_map<K, V, H>::iterator b, e;
        auto begin() const { return b; }
        auto end() const { return e; }
    };
    range_type range() { return range_type{this->begin(), this->end()}; }
    struct const_range_type {
        typename std::unordered_map<K, V, H>::const_iterator b, e;
        auto begin() const { return b; }
        auto end() const { return e; }
    };
    const_range_type range() const { return const_range_type{this->begin(), this->end()}; }
};

template <class K, class V, class H = std::hash<K>> class concurrent_unordered_multimap {
public:
    using value_type = std::pair<const K, V>;
    class iterator {
    public:
        iterator(concurrent_unordered_multimap *m, size_t g, size_t i) : m_(m), g_(g), i_(i) {}
        value_type &operator*() const { return *m_->groups_[g_][m_->groups_[g_].size() - 1 - i_]; }
        value_type *operator->() const { return &**this; }
        iterator &operator++() { if (++i_ == m_->groups_[g_].size()) { ++g_; i_ = 0; } return *this; }
        iterator operator++(int) { iterator t = *this; ++*this; return t; }
        bool operator==(const iterator &o) const { return g_ == o.g_ && i_ == o.i_; }
        bool operator!=(const iterator &o) const { return !(*this == o); }
    private:
        concurrent_unordered_multimap *m_;
        size_t g_, i_;
    };
    struct range_type {
        iterator b, e;
        iterator begin() const { return b; }
        iterator end() const { return e; }
    };
    void insert(const value_type &v) {
        auto it = index_.find(v.first);
        size_t g;
        if (it == index_.end()) { g = groups_.size(); index_.emplace(v.first, g); groups_.emplace_back(); }
        else g = it->second;
        store_.push_back(v);
        groups_[g].push_back(&store_.back());
    }
    size_t size() const { return store_.size(); }
    iterator begin() { return iterator(this, 0, 0); }
    iterator end() { return iterator(this, groups_.size(), 0); }
    range_type range() { return range_type{begin(), end()}; }
private:
    std::deque<value_type> store_;                      // stable addresses
    std::vector<std::vector<value_type *>> groups_;     // per key, in insertion order (iterated backwards)
    std::unordered_map<K, size_t, H> index_;
};

}  // namespace tbb
