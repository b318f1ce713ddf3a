// Stand-in for <boost/filesystem.hpp> (see ../README.md): std::filesystem under the boost name.
#pragma once
#include <climits>
#include <filesystem>
#include <fstream>
namespace boost { namespace filesystem {
using namespace std::filesystem;
class ifstream : public std::ifstream {
public:
    ifstream() = default;
    explicit ifstream(const std::filesystem::path &p) : std::ifstream(p) {}
    void open(const std::filesystem::path &p) { std::ifstream::open(p); }
};
class ofstream : public std::ofstream {
public:
    ofstream() = default;
    explicit ofstream(const std::filesystem::path &p) : std::ofstream(p) {}
    void open(const std::filesystem::path &p) { std::ofstream::open(p); }
};
}}  // namespace boost::filesystem
