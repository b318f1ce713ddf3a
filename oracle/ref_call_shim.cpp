/* ref_call_shim.cpp — TEST INFRASTRUCTURE (only tests/ may load the library built from this).
 *
 * The REFERENCE'S OWN function caller — /root/reference/src/call_functions.{h,tcc}, kept_kmer_db.h,
 * kmer_data.h (for_each_kmer), operators.h, fasta_parser.{h,cc}, unmodified and in place — compiled against
 * the stand-in third-party headers of oracle/refshim/ (Boost.Math statistics, Boost.Accumulators,
 * Boost.Regex and TBB restated there; see oracle/refshim/README.md) as FunctionCaller<KeptKmerDB<8>>, the
 * instantiation the reference's recall pass uses (src/kmers-build-signatures.cc:260-270).  What the product's
 * host/function_caller.h must reproduce (tests/test_reference_shim.py).
 *
 * Built by `make -C oracle ref` into oracle/_ref/libref_call.so when /root/reference is present.
 */
#include "signature_build.h"
#include "kept_kmer_db.h"
#include "call_functions.h"

#include <cstring>
#include <sstream>

extern "C" {

/* Calls for every record of a FASTA text against a kept table given as arrays (any row order).
 * Output: per record "id \t function \t function_index \t score \n", preceded — when want_calls — by one
 * "#call \t start \t end \t count \t function_index \t median \t mad" line per region call.
 * Returns the bytes needed. */
unsigned long long ref_call_functions(unsigned long long n_rows, const char *kmers, const uint16_t *avg_from_end,
                                      const uint16_t *function_index, const uint16_t *mean, const uint16_t *median,
                                      const uint16_t *var, const char *function_index_file, const char *fasta,
                                      unsigned long long fasta_len, int ignore_hypo, int want_calls, char *out,
                                      unsigned long long cap) {
    KeptKmers<8> kept;
    for (unsigned long long i = 0; i < n_rows; ++i) {
        Kmer<8> k;
        std::memcpy(k.data(), kmers + 8 * i, 8);
        kept.emplace(k, KeptKmer<8>{k, {avg_from_end[i], function_index[i], mean[i], median[i], var[i]}});
    }
    KeptKmerDB<8> db(kept);
    FunctionCaller<KeptKmerDB<8>> caller(db, fs::path(function_index_file));
    caller.ignore_hypothetical(ignore_hypo != 0);
    std::ostringstream buf;
    auto hit_cb = [](const std::string &, const Kmer<8> &, size_t, double, const StoredKmerData &) {};
    FastaParser parser;
    parser.set_callback([&](const std::string &id, const std::string &seq) {
        if (id.empty()) return 0;
        auto calls = std::make_shared<std::vector<KmerCall>>();
        caller.process_aa_seq(id, seq, calls, hit_cb);
        if (want_calls)
            for (const auto &c : *calls)
                buf << "#call\t" << c.start << "\t" << c.end << "\t" << c.count << "\t" << c.function_index << "\t"
                    << c.protein_length_median << "\t" << c.protein_length_med_avg_dev << "\n";
        FunctionIndex fi;
        std::string func;
        float score, offset;
        caller.find_best_call(id, *calls, fi, func, score, offset);
        buf << id << "\t" << func << "\t" << fi << "\t" << score << "\n";
        return 0;
    });
    std::istringstream in(std::string(fasta, fasta_len));
    parser.parse(in);
    parser.parse_complete();
    const std::string s = buf.str();
    if (s.size() <= cap) std::memcpy(out, s.data(), s.size());
    return s.size();
}

}
