// Stand-in for <nudb/nudb.hpp> (see ../README.md): just enough for src/nudb_kmer_db.h to compile.  The NuDB
// output of the reference command line is never requested by the tests, so these do nothing.
#pragma once
#include <cstddef>
#include <string>
#include <system_error>
namespace nudb {
using error_code = std::error_code;
struct xxhasher {};
inline unsigned long make_salt() { return 0; }
inline std::size_t block_size(const std::string &) { return 4096; }
template <class Hasher, class... Args> void create(Args &&...) {}
class store {
public:
    bool is_open() const { return false; }
    void close(error_code &) {}
    template <class... Args> void open(Args &&...) {}
    template <class... Args> void insert(Args &&...) {}
    template <class... Args> void fetch(Args &&...) {}
};
}  // namespace nudb
