// ref_fasta_shim.cpp — C entry point over the REFERENCE's own FastaParser
// (src/fasta_parser.{h,cc}, compiled where it lies by oracle/Makefile into
// oracle/_ref/libref_fasta.so).  TEST INFRASTRUCTURE ONLY: the tests use it to
// check that the drop-in's host-side FASTA reader delivers the same
// (id, def, seq) records on awkward inputs.  No reference source is copied.
#include "fasta_parser.h"

#include <cstdint>
#include <cstring>
#include <sstream>
#include <string>

extern "C" {

// Parses `len` bytes; writes records as id '\x01' def '\x01' seq '\x02' ... into out (capacity cap).
// Returns the number of bytes needed (call again with a larger buffer if > cap).
uint64_t ref_fasta_parse(const char *data, uint64_t len, char *out, uint64_t cap) {
    std::string buf;
    FastaParser parser;
    parser.set_def_callback([&buf](const std::string &id, const std::string &def, const std::string &seq) {
        buf += id; buf += '\x01'; buf += def; buf += '\x01'; buf += seq; buf += '\x02';
        return 0;
    });
    std::istringstream in(std::string(data, len));
    parser.parse(in);
    parser.parse_complete();      // src/signature_build.tcc:100-101 calls both
    if (buf.size() <= cap) std::memcpy(out, buf.data(), buf.size());
    return buf.size();
}

}
