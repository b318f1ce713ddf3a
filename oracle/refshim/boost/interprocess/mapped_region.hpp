// Stand-in for <boost/interprocess/mapped_region.hpp> (see ../../README.md).
#pragma once
#include "file_mapping.hpp"
namespace boost { namespace interprocess {
class mapped_region {
public:
    mapped_region() = default;
    mapped_region(const file_mapping &, mode_t) {}
    void *get_address() const { return nullptr; }
};
}}  // namespace boost::interprocess
