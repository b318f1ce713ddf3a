// host_capi.cpp — C exports of the host-side pieces for the CPU test-suite
// (no GPU, no libsigk calls): FASTA reader, SEED text helpers, function caller.
#include "function_caller.h"

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


// The function caller over a caller-supplied kept table (rows sorted by k-mer bytes).  function_names: one
// name per line, index = line number.  Writes, per FASTA record, "id \t function \t index \t score \n" and, when
// want_calls != 0, one "#call \t start \t end \t count \t function_index \t median \t mad" line per region call
// before it.  Returns the bytes needed.
uint64_t sigk_host_call_functions(uint64_t n_rows, const char *kmers, const uint16_t *avg_from_end, const uint16_t *function_index,
                                  const uint16_t *mean, const uint16_t *median, const uint16_t *var, const char *function_names,
                                  const char *fasta, uint64_t fasta_len, int ignore_hypo, int want_calls, char *out, uint64_t cap) {
    sigk_table t{};
    t.n_kept = n_rows; t.kmer = kmers; t.avg_from_end = avg_from_end; t.function_index = function_index;
    t.mean = mean; t.median = median; t.var = var;
    const SortedKmerDb db(t);
    std::vector<std::string> names;
    {
        std::istringstream in(function_names);
        std::string line;
        while (std::getline(in, line)) names.push_back(line);
    }
    FunctionCaller<SortedKmerDb> caller(db, names);
    caller.ignore_hypothetical(ignore_hypo != 0);
    std::ostringstream buf;
    FastaReader reader([&](const std::string &id, const std::string &, const std::string &seq) {
        if (id.empty()) return;
        std::vector<KmerCall> calls;
        auto hit_cb = [](const std::string &, const std::array<char, kCallK> &, size_t, double, const StoredKmerData &) {};
        caller.process_aa_seq(id, seq, &calls, hit_cb);
        if (want_calls)
            for (const auto &c : calls)
                buf << "#call\t" << c.start << "\t" << c.end << "\t" << c.count << "\t" << c.function_index << "\t"
                    << c.protein_length_median << "\t" << c.protein_length_med_avg_dev << "\n";
        const BestCall best = caller.find_best_call(calls);
        buf << id << "\t" << best.function << "\t" << best.function_index << "\t" << best.score << "\n";
    }, true);
    std::istringstream in(std::string(fasta, fasta_len));
    reader.parse(in);
    reader.finish();
    const std::string s = buf.str();
    if (s.size() <= cap) std::memcpy(out, s.data(), s.size());
    return s.size();
}

// windows for_each_kmer visits: offsets as u32; returns their number
uint64_t sigk_host_call_windows(const char *seq, uint64_t len, uint32_t *offsets, uint64_t cap) {
    uint64_t n = 0;
    for_each_kmer(std::string(seq, len), [&](const std::array<char, kCallK> &, size_t off) {
        if (n < cap) offsets[n] = (uint32_t)off;
        ++n;
    });
    return n;
}


// final.kmers of a caller-supplied table (the writer of the drop-in command line); returns 0 on success
int sigk_host_write_final_kmers(const char *path, uint64_t n_rows, const char *kmers, const uint16_t *avg_from_end,
                                const uint16_t *function_index, int n_threads) {
    sigk_table t{};
    t.n_kept = n_rows; t.kmer = kmers; t.avg_from_end = avg_from_end; t.function_index = function_index;
    return write_final_kmers(path, t, n_threads) ? 0 : 1;
}


// Everything the drop-in command line writes, given the kept table from outside (tests pass one in): the
// host phases (FunctionMap, gates) run as in kmers-build-signatures, then function.index, otu.index, genomes,
// final.kmers, distinct_functions and recall.report.d (host lookups) go to out_dir.  File lists in readdir order
// like the reference's populate_path_list.  Optional files may be "".  Returns 0 on success.
int sigk_host_outputs(const char *definition_dir, const char *fasta_dir, const char *good_functions_file, const char *good_roles_file,
                      const char *ignored_functions_file, const char *deleted_fids_file, int min_reps, int n_threads, const char *out_dir,
                      uint64_t n_rows, const char *kmers, const uint16_t *avg_from_end, const uint16_t *function_index,
                      const uint16_t *mean, const uint16_t *median, const uint16_t *var, const uint32_t *distinct_functions) {
    std::vector<fs::path> definitions, fasta;
    populate_path_list({definition_dir}, definitions);
    populate_path_list({fasta_dir}, fasta);
    std::vector<std::string> good_functions, good_roles;
    if (*good_functions_file) load_strings({good_functions_file}, good_functions);
    if (*good_roles_file) load_strings({good_roles_file}, good_roles);
    HostSignatureBuilder builder(n_threads, 100000);
    builder.load_function_data(good_functions, good_roles, definitions);
    const std::set<std::string> deleted = load_set_from_file(deleted_fids_file), ignored = load_set_from_file(ignored_functions_file);
    const fs::path dir = out_dir;
    ensure_directory(dir);
    builder.load_fasta(fasta, false, deleted);
    builder.process_kept_functions(min_reps, dir, ignored);
    std::ofstream(dir / "otu.index").close();
    { std::ofstream genomes(dir / "genomes"); genomes << "empty genomes\n"; }
    sigk_table t{};
    t.n_kept = n_rows; t.kmer = kmers; t.avg_from_end = avg_from_end; t.function_index = function_index;
    t.mean = mean; t.median = median; t.var = var; t.distinct_functions = distinct_functions;
    if (!write_final_kmers(dir / "final.kmers", t, n_threads)) return 1;
    write_distinct_functions(dir / "distinct_functions", t, builder.function_map());
    std::error_code ec;
    fs::create_directory(dir / "recall.report.d", ec);
    return write_recall_reports(builder.function_map(), builder.all_fasta_data(), t, dir / "function.index", dir / "recall.report.d",
                                n_threads, BatchLookup()) ? 0 : 1;
}

}
