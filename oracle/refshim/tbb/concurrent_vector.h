// Stand-in for <tbb/concurrent_vector.h> (see README.md).
#pragma once
#include <vector>
namespace tbb {
template <class T> class concurrent_vector : public std::vector<T> {
public:
    using std::vector<T>::vector;
    struct range_type {
        typename std::vector<T>::iterator b, e;
        auto begin() const { return b; }
        auto end() const { return e; }
    };
    range_type range() { return range_type{this->begin(), this->end()}; }
};
}  // namespace tbb
