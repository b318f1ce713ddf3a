// kmers-build-signatures — drop-in for the reference command line
// (src/kmers-build-signatures.cc): same options (:47-62), same input
// conventions, same files under --kmer-data-dir; the signature build itself
// (extract_kmers + process_kmers, :194-196) runs on the GPU through libsigk.
//
// Written: function.index, otu.index (empty), genomes, final.kmers
// (KMER \t avg_from_end \t function_index \t \n, :213-217), distinct_functions,
// the stdout counters.  Rows of final.kmers are in k-mer order (the reference's
// order is TBB hash order; write-cmph-from-kmers.cc:28-38 indexes by k-mer, so
// no consumer depends on it), recall.report.d/<fasta file> (:266-349, through
// function_caller.h).  --perfect-hash writes kmer_data.sigk (the sorted kept
// table, what this repo's kmers-call-functions opens) instead of a cmph hash;
// the NuDB output is not built (libraries absent).
//
// Extra options: --device N, --sorted-files (deterministic file order instead
// of readdir order), --dump-packed FILE (write the gated packed proteins and
// stop before the GPU: host-logic tests), --sigk-table FILE (write the table
// file without --perfect-hash), --no-recall (skip the recall pass),
// --host-recall (recall lookups on the host instead of sigk_lookup),
// --max-seqs-per-file N (the constant of :18, default 100000; tests),
// --gpu-fasta (both FASTA passes read their files through sigk_fasta_parse /
// sigk_fasta_commit instead of the host reader: same records, the proteins packed
// on the device), --timings (wall time of every phase of main on stderr).
#include "function_caller.h"

#include <atomic>
#include <chrono>
#include <cstring>
#include <mutex>
#include <thread>

using namespace sigk_host;

namespace {

struct Options {
    std::vector<std::string> definition_dirs, fasta_dirs, fasta_keep_dirs, good_function_files, good_role_files;
    fs::path deleted_fids_file, ignored_functions_file, kmer_data_dir, final_kmers, perfect_hash, perfect_hash_data, dump_packed, sigk_table;
    bool no_recall = false, host_recall = false, gpu_fasta = false, timings = false;
    int max_seqs_per_file = 100000;                                 // MaxSequencesPerFile, :18
    std::string nudb_file;
    int min_reps_required = 3, n_threads = 1, device = 0;
    bool sorted_files = false, help = false;
};

void usage(const char *argv0) {
    std::cout << "Usage: " << argv0 << " [options]\nAllowed options:\n"
              << "  -D [ --definition-dir ] arg          Directory of function definition files\n"
              << "  -F [ --fasta-dir ] arg               Directory of fasta files of protein data\n"
              << "  -K [ --fasta-keep-functions-dir ] arg Directory of fasta files of protein data (keep functions defined here)\n"
              << "  --good-functions arg                 File containing list of functions to be kept\n"
              << "  --good-roles arg                     File containing list of roles to be kept\n"
              << "  --deleted-features-file arg          File containing list of deleted feature IDs\n"
              << "  --ignored-functions-file arg         File containing list of functions for which we do not create signatures\n"
              << "  --kmer-data-dir arg                  Write kmer data files to this directory\n"
              << "  --nudb-file arg                      (accepted; NuDB output is not built)\n"
              << "  --min-reps-required arg              Minimum number of genomes a function must be seen in\n"
              << "  --final-kmers arg                    Write final.kmers file\n"
              << "  --n-threads arg                      (accepted; the build runs on the GPU)\n"
              << "  --perfect-hash arg / --perfect-hash-data arg  (accepted; cmph output is not built)\n"
              << "  --device arg / --sorted-files / --dump-packed arg / --sigk-table arg / --no-recall / --host-recall / --max-seqs-per-file arg\n"
              << "  --gpu-fasta (parse the FASTA files on the GPU: sigk_fasta_parse) / --timings (wall time of every phase on stderr)\n"
              << "  -h [ --help ]                        show this help message\n";
}

bool parse(int argc, char **argv, Options &o) {
    auto is_opt = [](const char *s) { return s[0] == '-' && s[1] != '\0'; };
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i], val;
        const size_t eq = a.find('=');
        bool has_val = false;
        if (a.rfind("--", 0) == 0 && eq != std::string::npos) { val = a.substr(eq + 1); a = a.substr(0, eq); has_val = true; }
        auto next = [&]() -> std::string {
            if (has_val) return val;
            if (i + 1 >= argc) { std::cerr << "option " << a << " needs a value\n"; std::exit(1); }
            return argv[++i];
        };
        auto multi = [&](std::vector<std::string> &dst) {            // boost multitoken: all following non-option words
            if (has_val) { dst.push_back(val); return; }
            if (i + 1 >= argc || is_opt(argv[i + 1])) { std::cerr << "option " << a << " needs a value\n"; std::exit(1); }
            while (i + 1 < argc && !is_opt(argv[i + 1])) dst.push_back(argv[++i]);
        };
        if (a == "-h" || a == "--help") o.help = true;
        else if (a == "-D" || a == "--definition-dir") multi(o.definition_dirs);
        else if (a == "-F" || a == "--fasta-dir") multi(o.fasta_dirs);
        else if (a == "-K" || a == "--fasta-keep-functions-dir") o.fasta_keep_dirs.push_back(next());
        else if (a == "--good-functions") o.good_function_files.push_back(next());
        else if (a == "--good-roles") o.good_role_files.push_back(next());
        else if (a == "--deleted-features-file") o.deleted_fids_file = next();
        else if (a == "--ignored-functions-file") o.ignored_functions_file = next();
        else if (a == "--kmer-data-dir") o.kmer_data_dir = next();
        else if (a == "--nudb-file") o.nudb_file = next();
        else if (a == "--min-reps-required") o.min_reps_required = std::stoi(next());
        else if (a == "--final-kmers") o.final_kmers = next();
        else if (a == "--n-threads") o.n_threads = std::stoi(next());
        else if (a == "--perfect-hash") o.perfect_hash = next();
        else if (a == "--perfect-hash-data") o.perfect_hash_data = next();
        else if (a == "--device") o.device = std::stoi(next());
        else if (a == "--sorted-files") o.sorted_files = true;
        else if (a == "--dump-packed") o.dump_packed = next();
        else if (a == "--sigk-table") o.sigk_table = next();
        else if (a == "--no-recall") o.no_recall = true;
        else if (a == "--host-recall") o.host_recall = true;
        else if (a == "--max-seqs-per-file") o.max_seqs_per_file = std::stoi(next());
        else if (a == "--gpu-fasta") o.gpu_fasta = true;
        else if (a == "--timings") o.timings = true;
        else { std::cerr << "unrecognised option '" << a << "'\n"; return false; }
    }
    return true;
}

}  // namespace

int main(int argc, char **argv) {
    Options o;
    if (!parse(argc, argv, o)) return 1;
    if (o.help) { usage(argv[0]); return 1; }                       // the reference returns 1 after --help (:155-158)

    std::vector<fs::path> function_definitions, fasta_data, fasta_keep;
    populate_path_list(o.definition_dirs, function_definitions, o.sorted_files);
    populate_path_list(o.fasta_dirs, fasta_data, o.sorted_files);
    populate_path_list(o.fasta_keep_dirs, fasta_keep, o.sorted_files);
    std::cout << "definitions: "; for (auto &x : o.definition_dirs) std::cout << x << " "; std::cout << std::endl;
    std::cout << "fasta: ";       for (auto &x : o.fasta_dirs) std::cout << x << " ";      std::cout << std::endl;
    std::cout << "keep: ";        for (auto &x : o.fasta_keep_dirs) std::cout << x << " "; std::cout << std::endl;
    std::vector<std::string> good_functions, good_roles;
    load_strings(o.good_function_files, good_functions);
    load_strings(o.good_role_files, good_roles);

    // --timings: where a run's wall time goes, phase by phase of the reference's main()
    auto t_last = std::chrono::steady_clock::now();
    auto phase = [&](const char *name) {
        const auto now = std::chrono::steady_clock::now();
        if (o.timings) std::cerr << "[time] " << name << ": " << std::chrono::duration<double>(now - t_last).count() << " s\n";
        t_last = now;
    };
    HostSignatureBuilder builder(o.n_threads, o.max_seqs_per_file);
    if (o.gpu_fasta && o.dump_packed.empty()) { if (builder.enable_gpu_fasta(o.device)) return 1; phase("create handle"); }
    builder.load_function_data(good_functions, good_roles, function_definitions);
    phase("load function data");
    const std::set<std::string> deleted_fids = load_set_from_file(o.deleted_fids_file);
    const std::set<std::string> ignored_functions = load_set_from_file(o.ignored_functions_file);
    ensure_directory(o.kmer_data_dir);

    std::cerr << "load fasta\n";
    builder.load_fasta(fasta_data, false, deleted_fids);
    builder.load_fasta(fasta_keep, true, deleted_fids);
    phase("load fasta (function evidence)");
    builder.process_kept_functions(o.min_reps_required, o.kmer_data_dir, ignored_functions);
    phase("process kept functions");
    if (!o.kmer_data_dir.empty()) {
        std::ofstream(o.kmer_data_dir / "otu.index").close();
        std::ofstream genomes(o.kmer_data_dir / "genomes");
        genomes << "empty genomes\n";
    }

    std::cerr << "extract kmers\n";
    if (o.gpu_fasta && o.dump_packed.empty()) { if (builder.extract_kmers_gpu(deleted_fids, o.device)) return 1; }
    else builder.extract_kmers(deleted_fids);
    phase("extract kmers (fasta -> packed proteins)");

    if (!o.dump_packed.empty()) {       // host-logic tests: the packed proteins libsigk would receive
        const sigk_proteins p = builder.packed();
        std::ofstream d(o.dump_packed, std::ios::binary);
        const uint64_t np = p.n_proteins, total = np ? p.starts[np] : 0;
        d.write((const char *)&np, 8); d.write((const char *)&total, 8);
        d.write((const char *)p.starts, (np + 1) * 8);
        d.write((const char *)p.function_index, np * 2);
        d.write((const char *)p.seq_id, np * 4);
        d.write((const char *)p.residues, total);
        std::cerr << "dumped " << np << " proteins, " << total << " residues\n";
        return 0;
    }

    std::cerr << "process kmers\n";
    sigk_table t;
    if (builder.process_kmers(o.device, &t)) return 1;
    phase("process kmers (GPU build incl. copies)");

    if (!o.final_kmers.empty()) {
        fs::path fk = o.final_kmers;
        if (fk.is_relative()) { fk = o.kmer_data_dir / fk; std::cerr << "Updated final_kmers to " << fk << "\n"; }
        std::cerr << "writing kmers to " << fk << "\n";
        if (!write_final_kmers(fk, t, o.n_threads)) { std::cerr << "error writing " << fk << "\n"; return 1; }
        std::cerr << "writing kmers to " << fk << " complete\n";
    }
    phase("write final.kmers");
    write_distinct_functions(o.kmer_data_dir / "distinct_functions", t, builder.function_map());
    const fs::path report_dir = o.kmer_data_dir / "recall.report.d";
    std::error_code ec;
    if (!fs::create_directory(report_dir, ec)) std::cerr << "mkdir " << report_dir << " failed\n";

    // the table in the form this repo's kmers-call-functions opens (stands in for the cmph pair, :256-265)
    if (!o.perfect_hash.empty() || !o.sigk_table.empty()) {
        fs::path tf = o.sigk_table.empty() ? fs::path("kmer_data.sigk") : o.sigk_table;
        if (tf.is_relative()) tf = o.kmer_data_dir / tf;
        if (!SortedKmerDb::write_file(tf, t)) { std::cerr << "cannot write " << tf << "\n"; return 1; }
        if (!o.perfect_hash.empty())
            std::cerr << "note: cmph is not available; wrote the sorted kept table to " << tf << " instead of a perfect hash\n";
    }
    if (!o.nudb_file.empty()) std::cerr << "note: the NuDB output is not built in this drop-in (library absent)\n";
    phase("write table file");

    // recall of the source data with the new k-mers (:266-349); the windows of a file are looked up in one batch on
    // the GPU unless --host-recall asks for host lookups (one handle: one caller at a time)
    if (!o.no_recall) {
        std::mutex lookup_mutex;
        BatchLookup lookup;
        if (!o.host_recall)
            lookup = [&](const uint8_t *res, const uint64_t *starts, uint64_t n, uint32_t *rows) {
                std::lock_guard<std::mutex> g(lookup_mutex);
                return builder.lookup(res, starts, n, rows);
            };
        std::cerr << "Begin recall\n";
        if (!write_recall_reports(builder.function_map(), builder.all_fasta_data(), t, o.kmer_data_dir / "function.index", report_dir,
                                  o.n_threads, lookup))
            return 1;
    }
    phase("recall reports");
    std::cerr << "all done\n";
    return 0;
}
