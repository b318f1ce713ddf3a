// kmers-call-functions — drop-in for the reference consumer CLI (src/kmers-call-functions.cc:47-197):
//
//   kmers-call-functions data-dir input-file [input-file ...]
//     -d/--data-dir DIR  -i/--input-files F...  -o/--output-files FILE  -j/--n-threads N
//     --ignore-hypo  --debug-hits  -h/--help
//
// Same positional arguments, options, output lines (`id \t function \t function_index \t score`) and
// stderr banner.  The database is <data-dir>/kmer_data.sigk (the sorted kept table written by this repo's
// kmers-build-signatures) instead of the cmph pair kmer_data.{mph,dat}: cmph is not available in this image,
// and a perfect hash adds nothing over a sorted table for correctness (the KmerDb concept only needs fetch,
// src/call_functions.h:60-66).  function.index is read from the data directory as in the reference.
// NOTE on semantics: the reference's CmphKmerDb::fetch (src/cmph_kmer.h:138-147) calls back for EVERY key,
// member or not — a minimal perfect hash maps a foreign k-mer to some unrelated slot — so its calls depend
// on cmph's seeded hash functions and cannot be reproduced without that library and that .mph file.  This
// tool uses exact membership, i.e. the semantics of KeptKmerDB (src/kept_kmer_db.h:20-28) that the
// reference's own recall pass runs with (src/kmers-build-signatures.cc:260-270).
// Extra option: --gpu DEVICE looks the windows of each input file up in one batch through libsigk
// (sigk_set_table + sigk_lookup); the calls are the same, only the lookups move.

#include "function_caller.h"

#include <atomic>
#include <mutex>
#include <thread>

using namespace sigk_host;

namespace {

struct Options {
    fs::path data_dir, output_file;
    std::vector<fs::path> input_files;
    bool debug_hits = false, ignore_hypo = false, help = false;
    int n_threads = 1;
    int gpu = -1;
};

void usage(const char *argv0) {
    std::cout << "Usage: " << argv0 << " data-dir input-file [input-file, ...]\nAllowed options:\n"
              << "  -d [ --data-dir ] arg      Data directory\n"
              << "  -i [ --input-files ] arg   Input files\n"
              << "  -o [ --output-files ] arg  Output file\n"
              << "  -j [ --n-threads ] arg     Number of threads\n"
              << "  --ignore-hypo              Ignore hypothetical protein kmers when making calls\n"
              << "  --debug-hits               Debug kmer hits\n"
              << "  --gpu arg                  (extra) batch the k-mer lookups on this CUDA device\n"
              << "  -h [ --help ]              show this help message\n\n";
}

bool parse(int argc, char **argv, Options &o) {
    std::vector<std::string> positional;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto value = [&](std::string &dst) {
            if (i + 1 >= argc) { std::cerr << "the required argument for option '" << a << "' is missing\n"; return false; }
            dst = argv[++i];
            return true;
        };
        std::string v;
        if (a == "-h" || a == "--help") o.help = true;
        else if (a == "--ignore-hypo") o.ignore_hypo = true;
        else if (a == "--debug-hits") o.debug_hits = true;
        else if (a == "-d" || a == "--data-dir") { if (!value(v)) return false; o.data_dir = v; }
        else if (a == "-o" || a == "--output-files") { if (!value(v)) return false; o.output_file = v; }
        else if (a == "-j" || a == "--n-threads") { if (!value(v)) return false; o.n_threads = std::stoi(v); }
        else if (a == "--gpu") { if (!value(v)) return false; o.gpu = std::stoi(v); }
        else if (a == "-i" || a == "--input-files") {           // multitoken: up to the next option
            while (i + 1 < argc && argv[i + 1][0] != '-') o.input_files.emplace_back(argv[++i]);
        } else if (!a.empty() && a[0] == '-') { std::cerr << "unrecognised option '" << a << "'\n"; return false; }
        else positional.push_back(a);
    }
    // positional: data-dir first (if not given by option), the rest are input files (:64-66)
    size_t k = 0;
    if (o.data_dir.empty() && k < positional.size()) o.data_dir = positional[k++];
    for (; k < positional.size(); ++k) o.input_files.emplace_back(positional[k]);
    return true;
}

}  // namespace

int main(int argc, char **argv) {
    Options o;
    if (!parse(argc, argv, o)) return 1;
    if (o.help) { usage(argv[0]); return 0; }
    if (o.input_files.empty()) { usage(argv[0]); return 1; }

    std::cerr << "Data size " << sizeof(StoredKmerData) << "\n";
    const fs::path db_base = o.data_dir / "kmer_data";
    SortedKmerDb db;
    if (!db.load_file(fs::path(db_base.string() + ".sigk"))) {
        std::cerr << "Database " << db_base << " does not exist\n";
        return 1;
    }
    FunctionCaller<SortedKmerDb> caller(db, o.data_dir / "function.index");
    caller.ignore_hypothetical(o.ignore_hypo);

    std::ofstream out_file;
    if (!o.output_file.empty()) out_file.open(o.output_file);
    std::ostream &anno_out = o.output_file.empty() ? std::cout : out_file;
    std::mutex out_mutex;

    auto hit_cb = [&](const std::string &, const std::array<char, kCallK> &kmer, size_t offset, double, const StoredKmerData &kd) {
        if (!o.debug_hits) return;
        std::ostringstream line;
        line.write(kmer.data(), kCallK);
        line << "\t" << offset << "\t" << caller.function_at_index(kd.function_index) << "\t" << kd.median << "\t" << kd.mean << "\t"
             << kd.var << "\t" << std::sqrt((double)kd.var) << "\t" << "\n";
        std::lock_guard<std::mutex> g(out_mutex);
        std::cout << line.str();
    };

    sigk_handle *gpu = nullptr;
    std::mutex gpu_mutex;
    if (o.gpu >= 0) {
        sigk_config cfg{SIGK_ABI_VERSION, SIGK_K, o.gpu, 0, 1, 0};
        sigk_table t{};
        t.n_kept = db.size();
        t.kmer = db.kmer_bytes();
        if (sigk_create(&cfg, &gpu)) { std::cerr << "sigk_create: " << sigk_last_error(nullptr) << "\n"; return 1; }
        if (sigk_set_table(gpu, &t)) { std::cerr << "libsigk: " << sigk_last_error(gpu) << "\n"; return 1; }
    }
    std::atomic<bool> failed{false};

    // one task per input file; every file's calls are written as one block (the reference's writer queue)
    std::atomic<size_t> next{0};
    auto worker = [&] {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= o.input_files.size()) break;
            std::ifstream in(o.input_files[i]);
            std::ostringstream buf;
            auto call_cb = [&](const std::string &id, const std::string &func, uint16_t fi, float score, size_t) {
                buf << id << "\t" << func << "\t" << fi << "\t" << score << "\n";
            };
            if (gpu) {
                auto lookup = [&](const uint8_t *res, const uint64_t *starts, uint64_t n, uint32_t *rows) {
                    std::lock_guard<std::mutex> g(gpu_mutex);            // one handle: one caller at a time
                    const int rc = sigk_lookup(gpu, res, starts, n, rows);
                    if (rc) std::cerr << "libsigk: " << sigk_last_error(gpu) << "\n";
                    return rc;
                };
                if (caller.process_fasta_stream_batched(in, lookup, hit_cb, call_cb)) failed = true;
            } else {
                caller.process_fasta_stream(in, hit_cb, call_cb);
            }
            const std::string text = buf.str();
            if (!text.empty()) {
                std::lock_guard<std::mutex> g(out_mutex);
                anno_out << text;
            }
        }
    };
    const int nt = std::max(1, std::min<int>(o.n_threads, (int)o.input_files.size()));
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
    anno_out.flush();
    if (gpu) sigk_destroy(gpu);
    return failed ? 1 : 0;
}
