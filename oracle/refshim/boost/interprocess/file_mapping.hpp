// Stand-in for <boost/interprocess/file_mapping.hpp> (see ../../README.md): only needed for src/cmph_kmer.h to
// compile; the tests never open a cmph database.
#pragma once
#include <sys/mman.h>
#include <cstring>
namespace boost { namespace interprocess {
enum mode_t { read_only, read_write };
class file_mapping {
public:
    file_mapping() = default;
    file_mapping(const char *, mode_t) {}
};
}}  // namespace boost::interprocess
