/* ref_signature_shim.cpp — TEST INFRASTRUCTURE (only tests/ may load the library built from this).
 *
 * Compiles the REFERENCE'S OWN signature builder — /root/reference/src/signature_build.{h,tcc},
 * function_map.h, seed_utils.h, kmer_data.h, fasta_parser.{h,cc}, unmodified and in place — against
 * the stand-in third-party headers of oracle/refshim/ (Boost and TBB are not in this image; see
 * oracle/refshim/README.md for what the stand-ins do and do not pin), and drives it the way
 * src/kmers-build-signatures.cc:163-196 does.  The kept table it produces is what oracle/sigk_oracle.cpp
 * must reproduce (tests/test_reference_shim.py).
 *
 * Built by `make -C oracle ref` into oracle/_ref/libref_signature.so when /root/reference is present.
 */
#include "signature_build.h"

#include <algorithm>
#include <cstdio>
#include <cstring>

namespace {

template <class List> void sorted_files(const char *dir, List &out) {
    std::vector<fs::path> v;
    if (dir && *dir)
        for (const auto &e : fs::directory_iterator(dir))
            if (fs::is_regular_file(e.path())) v.push_back(e.path());
    std::sort(v.begin(), v.end());          /* the reference takes readdir order; sorted here to be reproducible */
    for (auto &p : v) out.push_back(p);
}

}  // namespace

extern "C" {

/* Runs load_function_data / load_fasta / process_kept_functions / extract_kmers / process_kmers of
 * SignatureBuilder<8>.  Writes <out_dir>/function.index (by the reference's own writer) and
 * <out_dir>/ref_table.bin: "SIGKTBL1", u64 n, n x 8 k-mer bytes (sorted), then avg_from_end, function_index,
 * mean, median, var as u16 columns.  counters[0..2] = kept k-mers, distinct_signatures,
 * num_seqs_with_a_signature; distinct_functions / seqs_with_func: 65536 slots each.  Returns 0. */
static int g_max_seqs_per_file = 100000;      /* MaxSequencesPerFile, src/kmers-build-signatures.cc:18 */

/* tests only: sequence ids are file_number * max_seqs_per_file + n (src/signature_build.tcc:91), so a small value
 * makes the ids of different files collide without needing 100 000 proteins per file */
void ref_set_max_seqs_per_file(int v) { g_max_seqs_per_file = v; }

int ref_signature_build_ex(const char *definition_dir, const char *fasta_dir, const char *deleted_fids_file,
                           const char *good_functions_file, const char *good_roles_file, const char *ignored_functions_file,
                           int min_reps, int n_threads, const char *out_dir, unsigned long long *counters,
                           unsigned *distinct_functions, unsigned *seqs_with_func) {
    std::vector<fs::path> definitions, fasta;
    sorted_files(definition_dir, definitions);
    sorted_files(fasta_dir, fasta);
    /* one string per line, as load_strings / load_set_from_file of src/path_utils.h read them */
    auto lines_of = [](const char *file) {
        std::vector<std::string> v;
        if (file && *file) {
            std::ifstream in(file);
            std::string line;
            while (std::getline(in, line)) v.push_back(line);
        }
        return v;
    };
    std::set<std::string> deleted, ignored;
    for (auto &l : lines_of(deleted_fids_file)) deleted.insert(l);
    for (auto &l : lines_of(ignored_functions_file)) ignored.insert(l);
    SignatureBuilder<8> builder(n_threads, g_max_seqs_per_file);
    builder.load_function_data(lines_of(good_functions_file), lines_of(good_roles_file), definitions);
    builder.load_fasta(fasta, false, deleted);
    builder.process_kept_functions(min_reps, fs::path(out_dir), ignored);
    builder.extract_kmers(deleted);
    builder.process_kmers();

    struct Row { Kmer<8> k; StoredKmerData d; };
    std::vector<Row> rows;
    for (const auto &e : builder.kept_kmers()) rows.push_back(Row{e.first, e.second.stored_data});
    std::sort(rows.begin(), rows.end(), [](const Row &a, const Row &b) { return std::memcmp(a.k.data(), b.k.data(), 8) < 0; });
    const std::string path = (fs::path(out_dir) / "ref_table.bin").string();
    std::FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) return 1;
    const unsigned long long n = rows.size();
    std::fwrite("SIGKTBL1", 1, 8, f);
    std::fwrite(&n, 8, 1, f);
    for (const auto &r : rows) std::fwrite(r.k.data(), 1, 8, f);
    for (int c = 0; c < 5; ++c)
        for (const auto &r : rows) {
            const uint16_t v = c == 0 ? r.d.avg_from_end : c == 1 ? r.d.function_index : c == 2 ? r.d.mean : c == 3 ? r.d.median : r.d.var;
            std::fwrite(&v, 2, 1, f);
        }
    std::fclose(f);
    const KmerStatistics &st = builder.kmer_stats();
    counters[0] = n;
    counters[1] = (unsigned long long)(int)st.distinct_signatures;
    counters[2] = st.seqs_with_a_signature.size();
    std::memset(distinct_functions, 0, 65536 * sizeof(unsigned));
    std::memset(seqs_with_func, 0, 65536 * sizeof(unsigned));
    for (const auto &e : st.distinct_functions) distinct_functions[e.first & 0xFFFF] = (unsigned)e.second;
    for (const auto &e : st.seqs_with_func) seqs_with_func[e.first] = (unsigned)e.second;
    return 0;
}


int ref_signature_build(const char *definition_dir, const char *fasta_dir, const char *deleted_fids_file, int min_reps,
                        int n_threads, const char *out_dir, unsigned long long *counters, unsigned *distinct_functions,
                        unsigned *seqs_with_func) {
    return ref_signature_build_ex(definition_dir, fasta_dir, deleted_fids_file, "", "", "", min_reps, n_threads, out_dir, counters,
                                  distinct_functions, seqs_with_func);
}

}
