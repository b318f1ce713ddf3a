// Stand-in for <tbb/concurrent_unordered_set.h> (see README.md).
#pragma once
#include <unordered_set>
namespace tbb {
template <class T, class H = std::hash<T>> class concurrent_unordered_set : public std::unordered_set<T, H> {};
}  // namespace tbb
