// Stand-in for <tbb/concurrent_map.h> (see README.md).
#pragma once
#include <map>
namespace tbb {
template <class K, class V, class C = std::less<K>> class concurrent_map : public std::map<K, V, C> {};
}  // namespace tbb
