// Stand-in for <tbb/concurrent_vector.h> (see README.md).
#pragma once
#include <cassert>
#include <cstring>
#include <thread>
#include <vector>
namespace tbb {
template <class T> class concurrent_vector : public std::vector<T> {
public:
    using std::vector<T>::vector;
    // tbb's push_back returns an iterator to the new element
    typename std::vector<T>::iterator push_back(const T &v) { std::vector<T>::push_back(v); return this->end() - 1; }
    struct range_type {
        typename std::vector<T>::iterator b, e;
        auto begin() const { return b; }
        auto end() const { return e; }
    };
    range_type range() { return range_type{this->begin(), this->end()}; }
    struct const_range_type {
        typename std::vector<T>::const_iterator b, e;
        auto begin() const { return b; }
        auto end() const { return e; }
    };
    const_range_type range() const { return const_range_type{this->begin(), this->end()}; }
};
}  // namespace tbb
