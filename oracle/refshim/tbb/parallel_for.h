// Stand-in for <tbb/parallel_for.h> (see README.md): runs the body once over the whole range, serially.
#pragma once
#include <cstddef>
namespace tbb {
template <class T> class blocked_range {
public:
    blocked_range(T b, T e) : b_(b), e_(e) {}
    T begin() const { return b_; }
    T end() const { return e_; }
private:
    T b_, e_;
};
template <class Range, class Body> void parallel_for(const Range &r, const Body &body) { body(r); }
template <class Range, class Body> void parallel_for(Range &r, const Body &body) { body(r); }
}  // namespace tbb
