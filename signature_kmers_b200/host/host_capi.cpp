// host_capi.cpp — C exports of the host-side pieces for the CPU test-suite
// (no GPU, no libsigk calls): FASTA reader, SEED text helpers.
#include "signature_host.h"

#include <cstring>

using namespace sigk_host;

extern "C" {

// records as id \x01 def \x01 seq \x02 ...; returns the bytes needed
uint64_t sigk_host_fasta_parse(const char *data, uint64_t len, char *out, uint64_t cap) {
    std::string buf;
    FastaReader r([&](const std::string &id, const std::string &def, const std::string &seq) {
        buf += id; buf += '\x01'; buf += def; buf += '\x01'; buf += seq; buf += '\x02';
    }, true);
    std::istringstream in(std::string(data, len));
    r.parse(in);
    r.finish();
    if (buf.size() <= cap) std::memcpy(out, buf.data(), buf.size());
    return buf.size();
}

// func \x01 sep \x01 comment
uint64_t sigk_host_split_func_comment(const char *s, char *out, uint64_t cap) {
    std::string f, sep, c;
    split_func_comment(s, f, sep, c);
    const std::string buf = f + '\x01' + sep + '\x01' + c;
    if (buf.size() < cap) std::memcpy(out, buf.c_str(), buf.size() + 1);
    return buf.size();
}

uint64_t sigk_host_roles(const char *s, char *out, uint64_t cap) {
    std::string buf;
    for (const auto &r : roles_of_function(s)) { buf += r; buf += '\x01'; }
    if (buf.size() < cap) std::memcpy(out, buf.c_str(), buf.size() + 1);
    return buf.size();
}

int sigk_host_is_truncated(const char *s) { return is_truncated_comment(s) ? 1 : 0; }

}
