// signature_host.h — host side of the drop-in kmers-build-signatures: the
// reference's input conventions restated in plain C++17 (no Boost, no TBB).
//
//   FastaReader          src/fasta_parser.{h,cc}  (same state machine, same quirks)
//   FunctionMap          src/function_map.h:62-411
//   path helpers         src/path_utils.h:17-101
//   HostSignatureBuilder src/signature_build.{h,tcc}: the SignatureBuilder<K> API that
//                        kmers-build-signatures.cc calls; extract_kmers packs the gated
//                        proteins (tcc:83-160), process_kmers hands them to libsigk.
#pragma once

#include "../../include/sigk.h"
#include "seed_text.h"

#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <algorithm>
#include <atomic>
#include <set>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace sigk_host {

namespace fs = std::filesystem;

// ---------------------------------------------------------------------------
// FASTA: one record per '>' header; id = up to the first blank, def = the rest
// of the header line INCLUDING its leading blank, seq = letters and '*'.
// Quirks kept: CR dropped; '*' accepted mid-line but a line that
// STARTS with a non-letter is reported and the character dropped; a blank
// line keeps the record open; after a header-only record the next '>' is seen in
// the "data" state and is reported as a bad character, so that header's letters
// fall into the open record's sequence; the callback fires once more at end of
// input with whatever is pending (an empty id when nothing is).
class FastaReader {
public:
    using Callback = std::function<void(const std::string &id, const std::string &def, const std::string &seq)>;
    explicit FastaReader(Callback cb, bool quiet = false) : cb_(std::move(cb)), quiet_(quiet) {}

    // Same state machine, but runs of ordinary characters inside a sequence line, a definition or an id are
    // appended in one go; every other character takes the one-character step below.
    void feed(const char *p, size_t n) {
        static const CharClasses cc;
        const char *end = p + n;
        while (p < end) {
            const char *q = p;
            if (st_ == DATA) {
                while (q < end && cc.seq[(unsigned char)*q]) ++q;
                if (q > p) { seq_.append(p, q); p = q; continue; }
            } else if (st_ == DEF) {
                while (q < end && *q != '\n' && *q != '\r') ++q;
                if (q > p) { def_.append(p, q); p = q; continue; }
            } else if (st_ == ID) {
                while (q < end && !cc.id_stop[(unsigned char)*q]) ++q;
                if (q > p) { id_.append(p, q); p = q; continue; }
            }
            step(*p++);
        }
    }
    void parse(std::istream &in) {
        reset();
        char buf[1 << 16];
        while (in.read(buf, sizeof buf) || in.gcount()) feed(buf, (size_t)in.gcount());
        finish();
    }
    // parse() of the reference ends with parse_complete(), and its callers call it again
    void finish() { emit(); }

private:
    struct CharClasses {
        bool seq[256], id_stop[256];
        CharClasses() {
            for (int c = 0; c < 256; ++c) {
                seq[c] = std::isalpha(c) || c == '*';                           // what the DATA state keeps
                id_stop[c] = c == ' ' || c == '\t' || c == '\n' || c == '\r';   // what ends (or is skipped inside) an id
            }
        }
    };
    enum State { START, ID, DEF, DATA, LINE_START } st_ = START;
    std::string id_, def_, seq_;
    int line_ = 1;
    Callback cb_;
    bool quiet_;

    void reset() { st_ = START; id_.clear(); def_.clear(); seq_.clear(); }
    void emit() { cb_(id_, def_, seq_); id_.clear(); def_.clear(); seq_.clear(); }
    void complain(const std::string &what) {
        if (!quiet_) std::cerr << "Error found: " << what << " at line " << line_ << " id='" << id_ << "'" << std::endl;
    }
    void step(char c) {
        if (c == '\n') ++line_;
        if (c == '\r') return;
        switch (st_) {
        case START:
            if (c == '>') st_ = ID; else complain("Missing >");
            break;
        case ID:
            if (c == ' ' || c == '\t') { def_.push_back(c); st_ = DEF; }
            else if (c == '\n') st_ = DATA;
            else id_.push_back(c);
            break;
        case DEF:
            if (c == '\n') st_ = DATA; else def_.push_back(c);
            break;
        case DATA:
            if (c == '\n') st_ = LINE_START;
            else if (std::isalpha((unsigned char)c) || c == '*') seq_.push_back(c);
            else complain(std::string("Bad data character '") + c + "'");
            break;
        case LINE_START:
            if (c == '>') { emit(); st_ = ID; }
            else if (c == '\n') {}
            else if (std::isalpha((unsigned char)c)) { seq_.push_back(c); st_ = DATA; }
            else complain(std::string("Bad id or data character '") + c + "'");
            break;
        }
    }
};

// ---------------------------------------------------------------------------
// src/path_utils.h
inline void populate_path_list(const std::vector<std::string> &dirs, std::vector<fs::path> &paths, bool sorted = false) {
    for (const auto &dir : dirs) {
        std::vector<fs::path> here;
        for (const auto &ent : fs::directory_iterator(dir))        // readdir order, like boost::filesystem (:17-31)
            if (fs::is_regular_file(ent.path())) here.push_back(ent.path());
        if (sorted) std::sort(here.begin(), here.end());
        paths.insert(paths.end(), here.begin(), here.end());
    }
}
inline void load_strings(const std::vector<std::string> &files, std::vector<std::string> &out) {
    for (const auto &f : files) {
        std::ifstream in(f);
        if (!in.good()) { std::cerr << "could not open " << f << "\n"; continue; }
        std::string line;
        while (std::getline(in, line, '\n')) out.push_back(line);
    }
}
inline std::set<std::string> load_set_from_file(const fs::path &file) {
    std::set<std::string> s;
    if (!file.empty()) {
        std::ifstream in(file);
        std::string line;
        while (std::getline(in, line, '\n')) s.insert(line);
    }
    return s;
}
inline void ensure_directory(const fs::path &dir) {
    if (!dir.empty() && !fs::is_directory(dir) && !fs::create_directory(dir)) {
        std::cerr << "Error creating " << dir << "\n";
        std::exit(1);
    }
}

// ---------------------------------------------------------------------------
// accumulator_set<float, stats<mean, median, variance, count>> of function_map.h:463 —
// the per-function length statistics written to function.index (columns 3-7).  Same
// algorithms as the hot path's accumulator but with float state (sample type float).
// fn(i) for i in [0, n) on up to `threads` threads (work handed out one index at a time)
template <class Fn>
inline void parallel_for_index(size_t n, int threads, Fn fn) {
    const int nt = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(1, threads), n));
    if (nt <= 1) { for (size_t i = 0; i < n; ++i) fn(i); return; }
    std::atomic<size_t> next{0};
    auto worker = [&] { for (size_t i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i); };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
}

struct FloatStats {
    size_t n = 0;
    float sum = 0, var = 0;
    float q[5] = {0, 0, 0, 0, 0}, pos[5] = {1, 2, 3, 4, 5}, des[5] = {1, 2, 3, 4, 5};
    void push(double x) {
        static const float inc[5] = {0.f, 0.25f, 0.5f, 0.75f, 1.f};
        ++n;
        sum = (float)(sum + x);
        if (n <= 5) {
            q[n - 1] = (float)x;
            if (n == 5) std::sort(q, q + 5);
        } else {
            size_t k;
            if (x < q[0]) { q[0] = (float)x; k = 1; }
            else if (q[4] <= x) { q[4] = (float)x; k = 4; }
            else k = (size_t)(std::upper_bound(q, q + 5, x, [](double a, float b) { return a < b; }) - q);
            for (size_t i = k; i < 5; ++i) pos[i] += 1.f;
            for (size_t i = 0; i < 5; ++i) des[i] += inc[i];
            for (size_t i = 1; i <= 3; ++i) {
                const float d = des[i] - pos[i], dp = pos[i + 1] - pos[i], dm = pos[i - 1] - pos[i];
                const float hp = (q[i + 1] - q[i]) / dp, hm = (q[i - 1] - q[i]) / dm;
                if ((d >= 1.f && dp > 1.f) || (d <= -1.f && dm < -1.f)) {
                    const short s = (short)(d / std::fabs(d));
                    const float h = q[i] + s / (dp - dm) * ((s - dm) * hp + (dp - s) * hm);
                    if (q[i - 1] < h && h < q[i + 1]) q[i] = h;
                    else { if (d > 0) q[i] += hp; if (d < 0) q[i] -= hm; }
                    pos[i] += s;
                }
            }
        }
        if (n > 1) {
            const float tmp = (float)(x - (double)(sum / (float)n));
            var = (var * (float)(n - 1)) / (float)n + (tmp * tmp) / (float)(n - 1);
        }
    }
    // no samples: 0.f / 0 like the accumulator's sum / count (the reference's function.index shows "-nan" for
    // the always-present "hypothetical protein" row when no protein carries it)
    double mean() const { return sum / (float)n; }
    double median() const { return q[2]; }
};

// ---------------------------------------------------------------------------
// src/function_map.h
class FunctionMap {
public:
    void add_good_roles(const std::vector<std::string> &r) { good_roles_.insert(r.begin(), r.end()); }
    void add_good_functions(const std::vector<std::string> &r) { good_functions_.insert(r.begin(), r.end()); }

    // :62-104.  Reading and splitting the lines needs nothing from the map, so it can run for several files at
    // once (parse_id_assignments); the map updates keep the reference's order (apply_id_assignments).
    struct Assignment { std::string id, func, stripped; bool truncated; };
    static std::vector<Assignment> parse_id_assignments(const fs::path &file) {
        std::vector<Assignment> out;
        std::ifstream in(file);
        std::string line;
        int lineno = 0;
        while (std::getline(in, line)) {
            ++lineno;
            const size_t s = line.find('\t');
            if (s == std::string::npos) { std::cerr << "bad line " << lineno << " in file " << file << "\n"; continue; }
            const size_t s2 = line.find('\t', s + 1);
            Assignment a;
            a.id = line.substr(0, s);
            a.func = s2 == std::string::npos ? line.substr(s + 1) : line.substr(s + 1, s2 - s - 1);
            std::string delim, comment;
            split_func_comment(a.func, a.stripped, delim, comment);
            a.truncated = delim == "#" && is_truncated_comment(comment);
            out.push_back(std::move(a));
        }
        return out;
    }
    void apply_id_assignments(std::vector<Assignment> &rows) {
        by_id_.reserve(by_id_.size() + rows.size());
        for (auto &a : rows) {
            IdEntry &e = by_id_[a.id];
            e.has_original = true;
            if (!a.truncated) e.function = a.stripped;                       // a truncated one keeps any earlier assignment (:93-98)
            e.original_stripped = std::move(a.stripped);
            e.original = std::move(a.func);
        }
    }
    void load_id_assignments(const fs::path &file) {
        auto rows = parse_id_assignments(file);
        apply_id_assignments(rows);
    }

    // :120-238 — genome / function evidence from one FASTA file.  The parse (parse_fasta_headers: id, definition
    // and length of every record) is independent of the map; the evidence is applied in file and record order.
    struct FastaHeader { std::string id, def; size_t length; };
    static std::vector<FastaHeader> parse_fasta_headers(const fs::path &file) {
        std::vector<FastaHeader> out;
        std::ifstream in(file);
        FastaReader reader([&](const std::string &id, const std::string &def, const std::string &seq) {
            out.push_back(FastaHeader{id, def, seq.length()});
        });
        reader.parse(in);
        reader.finish();
        return out;
    }
    void apply_fasta_headers(const fs::path &file, const std::vector<FastaHeader> &records, bool keep_function_flag,
                             const std::set<std::string> &deleted_fids) {
        std::string genome;
        for (const auto &rec : records) {
            const std::string &id = rec.id, &def = rec.def;
            if (id.empty() || deleted_fids.count(id)) continue;
            std::string func;
            if (!def.empty()) {
                const size_t x = def.find_first_not_of(" \t");
                func = def.substr(x);           // npos -> throws like the reference would; defs are never all blank in practice
            }
            std::string genome_loc, m1;
            if (match_genome_def(def, m1, genome_loc)) {        // "\s+(.*)\s+\[([^]]+)\]$"
                std::string delim, comment;
                split_func_comment(m1, func, delim, comment);
                if (delim == "#" && is_truncated_comment(comment)) continue;
            } else genome_loc.clear();
            if (genome.empty()) {
                if (def.empty()) { std::string g; if (find_fig_genome(id, g)) genome = g; }
                else if (!genome_loc.empty()) genome = genome_loc;
            }
            if (genome.empty()) {
                genome = file.filename().string();
                if (!is_genome_id(genome)) std::cerr << "cannot determine genome from file " << file << "\n";
            }
            std::string &slot = by_id_[id].function;            // operator[]: inserts an empty entry like the reference
            if (slot.empty()) { if (!func.empty()) slot = func; }
            else func = slot;
            if (!func.empty()) {
                FunctionEvidence &ev = by_function_[func];
                ev.genomes.insert(genome);
                if (keep_function_flag) good_functions_.insert(func);
                ev.lengths.push((double)rec.length);
            }
        }
    }
    void load_fasta_file(const fs::path &file, bool keep_function_flag, const std::set<std::string> &deleted_fids) {
        apply_fasta_headers(file, parse_fasta_headers(file), keep_function_flag, deleted_fids);
    }

    // :257-332
    void process_kept_functions(int min_reps_required, const std::set<std::string> &ignored) {
        std::set<std::string> kept;
        for (const auto &entry : by_function_) {                 // any order: `kept` is a std::set, as in the reference
            const std::string &function = entry.first;
            bool ok = (int)entry.second.genomes.size() >= min_reps_required || good_functions_.count(function);
            if (!ok)
                for (const auto &role : roles_of_function(function))
                    if (good_roles_.count(role)) { ok = true; break; }
            if (ok) kept.insert(function);
        }
        kept.insert("hypothetical protein");
        for (const auto &fn : ignored) { std::cerr << "Ignore '" << fn << "'\n"; kept.erase(fn); }
        unsigned short next = 0;                                  // wraps at 65536 like the reference (:324-330)
        for (const auto &f : kept) {
            const unsigned short id = next++;
            function_index_map_[f] = id;
            index_function_map_[id] = f;
        }
        std::cout << "kept " << next << " functions\n";
    }

    std::string lookup_function(const std::string &id) const {
        auto it = by_id_.find(id);
        return it == by_id_.end() ? std::string() : it->second.function;
    }
    // src/function_map.h:351-358 (outputs untouched when the id was never assigned)
    void lookup_original_assignment(const std::string &id, std::string &func, std::string &stripped) const {
        auto it = by_id_.find(id);
        if (it != by_id_.end() && it->second.has_original) { func = it->second.original; stripped = it->second.original_stripped; }
    }
    std::string lookup_function(uint16_t idx) const {
        auto it = index_function_map_.find(idx);
        return it == index_function_map_.end() ? std::string() : it->second;
    }
    uint16_t lookup_index(const std::string &func) const {
        auto it = function_index_map_.find(func);
        return it == function_index_map_.end() ? (uint16_t)0xFFFF : it->second;
    }

    // :389-411  idx \t function \t count \t mean \t median \t var \t dev
    void write_function_index(const fs::path &dir) {
        std::ofstream of(dir / "function.index");
        std::map<int, std::string> by_index;
        for (const auto &e : function_index_map_) by_index.insert({e.second, e.first});
        for (const auto &e : by_index) {
            static const FloatStats none;
            auto st = by_function_.find(e.second);
            const FloatStats &a = st == by_function_.end() ? none : st->second.lengths;
            const double mean = a.mean(), median = a.median(), var = a.var;
            of << e.first << "\t" << e.second << "\t" << (int)a.n << "\t" << mean << "\t" << median << "\t" << var << "\t"
               << std::sqrt(var) << "\n";
        }
    }

private:
    // regex_match(def, "\\s+(.*)\\s+\\[([^]]+)\\]$"): leading blanks (all of them, greedy), greedy (.*),
    // one blank, "[genome]" at the very end; the greedy (.*) makes the rightmost usable '[' win.
    static bool match_genome_def(const std::string &def, std::string &m1, std::string &genome) {
        const size_t n = def.size();
        if (n < 4 || def[n - 1] != ']') return false;
        size_t a = 0;
        while (a < n && is_space(def[a])) ++a;
        if (a == 0) return false;
        for (size_t lb = n - 2; lb >= 1; --lb) {
            if (def[lb] == ']') return false;                     // the bracket body may not hold a ']'
            if (def[lb] != '[') continue;
            if (lb + 1 == n - 1) continue;                        // empty body: [^]]+ needs a character
            if (!is_space(def[lb - 1])) continue;
            if (a >= lb) {                                        // nothing but blanks before '[': two \s+ need two blanks
                if (lb < 2) return false;
                m1.clear();
            } else m1 = def.substr(a, lb - 1 - a);
            genome = def.substr(lb + 1, n - lb - 2);
            return true;
        }
        return false;
    }
    static bool find_fig_genome(const std::string &id, std::string &g) {      // "fig\|(\d+\.\d+)"
        for (size_t p = id.find("fig|"); p != std::string::npos; p = id.find("fig|", p + 1)) {
            size_t a = p + 4, b = a;
            while (b < id.size() && std::isdigit((unsigned char)id[b])) ++b;
            if (b == a || b >= id.size() || id[b] != '.') continue;
            size_t c = b + 1, d = c;
            while (d < id.size() && std::isdigit((unsigned char)id[d])) ++d;
            if (d == c) continue;
            g = id.substr(a, d - a);
            return true;
        }
        return false;
    }
    static bool is_genome_id(const std::string &s) {                          // "\d+\.\d+" full match
        size_t b = 0;
        while (b < s.size() && std::isdigit((unsigned char)s[b])) ++b;
        if (b == 0 || b >= s.size() || s[b] != '.') return false;
        size_t d = b + 1;
        while (d < s.size() && std::isdigit((unsigned char)s[d])) ++d;
        return d > b + 1 && d == s.size();
    }

    // the reference's id_function_map_, original_assignment_ and original_assignment_stripped_ (keyed by id, only ever
    // looked up) in one hash table; function_genome_map_ and function_accumulators_ (keyed by function) in another
    struct IdEntry { std::string function, original, original_stripped; bool has_original = false; };
    struct FunctionEvidence { std::set<std::string> genomes; FloatStats lengths; };
    std::unordered_map<std::string, IdEntry> by_id_;
    std::unordered_map<std::string, FunctionEvidence> by_function_;
    std::map<std::string, uint16_t> function_index_map_;
    std::map<uint16_t, std::string> index_function_map_;
    std::set<std::string> good_roles_, good_functions_;
};

// ---------------------------------------------------------------------------
// final.kmers, src/kmers-build-signatures.cc:206-221: KMER \t avg_from_end \t function_index \t \n per kept k-mer.
// Hundreds of millions of rows: blocks of rows are formatted on n_threads threads (at most 8 + 1 + 5 + 1 + 5 + 2
// bytes a row) and written in order.
inline bool write_final_kmers(const fs::path &file, const sigk_table &t, int n_threads) {
    std::FILE *f = std::fopen(file.c_str(), "w");
    if (!f) return false;
    const uint64_t block = 1 << 20;
    const uint64_t n_blocks = (t.n_kept + block - 1) / block;
    const size_t wave = (size_t)std::max(1, n_threads) * 2;
    auto put_u16 = [](char *p, unsigned v) {
        char tmp[5];
        int n = 0;
        do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
        while (n) *p++ = tmp[--n];
        return p;
    };
    std::vector<std::vector<char>> text(wave);
    bool ok = true;
    for (uint64_t b0 = 0; b0 < n_blocks && ok; b0 += wave) {
        const size_t nb = (size_t)std::min<uint64_t>(wave, n_blocks - b0);
        parallel_for_index(nb, n_threads, [&](size_t k) {
            const uint64_t lo = (b0 + k) * block, hi = std::min<uint64_t>(t.n_kept, lo + block);
            std::vector<char> &out = text[k];
            out.resize((size_t)(hi - lo) * 22);
            char *p = out.data();
            for (uint64_t i = lo; i < hi; ++i) {
                std::memcpy(p, t.kmer + 8 * i, 8); p += 8;
                *p++ = '\t'; p = put_u16(p, t.avg_from_end[i]);
                *p++ = '\t'; p = put_u16(p, t.function_index[i]);
                *p++ = '\t'; *p++ = '\n';
            }
            out.resize((size_t)(p - out.data()));
        });
        for (size_t k = 0; k < nb && ok; ++k) ok = std::fwrite(text[k].data(), 1, text[k].size(), f) == text[k].size();
    }
    return std::fclose(f) == 0 && ok;
}

// ---------------------------------------------------------------------------
// src/signature_build.{h,tcc}: same public calls, in the order main() makes them.
class HostSignatureBuilder {
public:
    HostSignatureBuilder(int n_threads, int max_seqs_per_file) : n_threads_(n_threads), max_seqs_per_file_(max_seqs_per_file) {}

    // Files are read and tokenised on n_threads threads; everything that touches the FunctionMap is applied in
    // the reference's order (file by file, record by record), so the outputs do not depend on the thread count.
    void load_function_data(const std::vector<std::string> &good_functions, const std::vector<std::string> &good_roles,
                            const std::vector<fs::path> &defs) {
        fm_.add_good_roles(good_roles);
        fm_.add_good_functions(good_functions);
        std::vector<std::vector<FunctionMap::Assignment>> parsed(defs.size());
        parallel_for_index(defs.size(), n_threads_, [&](size_t i) { parsed[i] = FunctionMap::parse_id_assignments(defs[i]); });
        for (auto &rows : parsed) fm_.apply_id_assignments(rows);
    }
    // --gpu-fasta: both FASTA passes (this one for the function evidence, extract_kmers_gpu for the proteins) read
    // their files through sigk_fasta_parse.  Creates the handle.
    int enable_gpu_fasta(int device) {
        gpu_fasta_ = true;
        return ensure_handle(device);
    }
    void load_fasta(const std::vector<fs::path> &files, bool /*keep_functions: dropped by the reference, tcc:32*/,
                    const std::set<std::string> &deleted) {
        std::vector<std::vector<FunctionMap::FastaHeader>> parsed(files.size());
        bool done = false;
        if (gpu_fasta_ && !files.empty()) {
            auto batch = std::make_unique<GpuFastaBatch>();
            if (gpu_parse(files, *batch) == 0) {
                report_fasta_errors(*batch);
                parallel_for_index(files.size(), n_threads_, [&](size_t i) {
                    auto &out = parsed[i];
                    const auto range = batch->records_of(i);
                    for (uint64_t r = range.first; r < range.second; ++r)
                        out.push_back(FunctionMap::FastaHeader{batch->id_of(r, i), batch->def_of(r, i), (size_t)(batch->seq_begin[r + 1] - batch->seq_begin[r])});
                    // the reference's callback fires once more per file for whatever is pending (the last record, already
                    // listed) and then again with an empty record; with no record at all both calls see an empty one
                    out.push_back(FunctionMap::FastaHeader{"", "", 0});
                    if (range.first == range.second) out.push_back(FunctionMap::FastaHeader{"", "", 0});
                });
                batch_ = std::move(batch);
                done = true;
            }
        }
        if (!done) parallel_for_index(files.size(), n_threads_, [&](size_t i) { parsed[i] = FunctionMap::parse_fasta_headers(files[i]); });
        for (size_t i = 0; i < files.size(); ++i) {
            fm_.apply_fasta_headers(files[i], parsed[i], false, deleted);
            all_fasta_data_.push_back(files[i]);
        }
    }
    void process_kept_functions(int min_reps, const fs::path &out_dir, const std::set<std::string> &ignored) {
        fm_.process_kept_functions(min_reps, ignored);
        if (!out_dir.empty()) fm_.write_function_index(out_dir);
    }

    // tcc:47-160 with the window loop removed: the gated proteins are packed in canonical order (file order,
    // record order inside a file).  Files are independent (tcc:58-68 runs them under tbb::parallel_for).
    void extract_kmers(const std::set<std::string> &deleted) {
        struct PerFile { std::vector<uint8_t> residues; std::vector<uint64_t> lengths; std::vector<uint16_t> func; std::vector<uint32_t> seq_id; };
        std::vector<PerFile> parts(all_fasta_data_.size());
        parallel_for_index(all_fasta_data_.size(), n_threads_, [&](size_t i) {
            PerFile &out = parts[i];
            std::ifstream in(all_fasta_data_[i]);
            unsigned next_sequence_id = (unsigned)i * (unsigned)max_seqs_per_file_;                 // :91
            FastaReader reader([&](const std::string &id, const std::string &, const std::string &seq) {
                if (deleted.count(id)) return;                                              // :94
                if (id.empty()) return;                                                     // :122
                const std::string func = fm_.lookup_function(id);
                if (func.empty()) return;                                                   // :133
                const unsigned seq_id = next_sequence_id++;                                 // :138
                const uint16_t fi = fm_.lookup_index(func);
                if (fi == 0xFFFF) return;                                                   // :155
                out.residues.insert(out.residues.end(), seq.begin(), seq.end());
                out.lengths.push_back(seq.size());
                out.func.push_back(fi);
                out.seq_id.push_back(seq_id);
            }, true);
            reader.parse(in);
            reader.finish();
        });
        residues_.clear(); starts_.assign(1, 0); func_.clear(); seq_id_.clear();
        size_t total = 0, count = 0;
        for (const auto &part : parts) { total += part.residues.size(); count += part.func.size(); }
        residues_.reserve(total); starts_.reserve(count + 1); func_.reserve(count); seq_id_.reserve(count);
        for (const auto &part : parts) {
            residues_.insert(residues_.end(), part.residues.begin(), part.residues.end());
            for (uint64_t len : part.lengths) starts_.push_back(starts_.back() + len);
            func_.insert(func_.end(), part.func.begin(), part.func.end());
            seq_id_.insert(seq_id_.end(), part.seq_id.begin(), part.seq_id.end());
        }
    }

    // The same gating with the files parsed on the GPU (sigk_fasta_parse, csrc/fasta.cu): the files' bytes go to the
    // device once, the parser's record table comes back, the ids are looked up here (FunctionMap is a string-keyed
    // map), and sigk_fasta_commit packs the kept records' residues on the device — process_kmers then builds from
    // what is already there.  Files stay independent: a file's records are a contiguous run of the table.
    int extract_kmers_gpu(const std::set<std::string> &deleted, int device) {
        if (int rc = ensure_handle(device)) return rc;
        // the parse of load_fasta is still the handle's latest one when it covered exactly these files
        if (!batch_ || batch_->files != all_fasta_data_) {
            batch_ = std::make_unique<GpuFastaBatch>();
            if (int rc = gpu_parse(all_fasta_data_, *batch_)) return rc;
        }
        const GpuFastaBatch &bt = *batch_;
        std::vector<uint8_t> keep(bt.n_records, 0);
        std::vector<uint16_t> func(bt.n_records, 0);
        std::vector<uint32_t> seq_id(bt.n_records, 0);
        parallel_for_index(bt.files.size(), n_threads_, [&](size_t i) {
            unsigned next_sequence_id = (unsigned)i * (unsigned)max_seqs_per_file_;                 // :91
            const auto range = bt.records_of(i);
            for (uint64_t r = range.first; r < range.second; ++r) {
                const std::string id = bt.id_of(r, i);
                if (deleted.count(id)) continue;                                            // :94
                if (id.empty()) continue;                                                   // :122
                const std::string fn = fm_.lookup_function(id);
                if (fn.empty()) continue;                                                   // :133
                const unsigned sid = next_sequence_id++;                                    // :138
                const uint16_t fi = fm_.lookup_index(fn);
                if (fi == 0xFFFF) continue;                                                 // :155
                keep[r] = 1; func[r] = fi; seq_id[r] = sid;
            }
        });
        const int rc = sigk_fasta_commit(h_, keep.data(), func.data(), seq_id.data());
        batch_.reset();
        if (rc) { std::cerr << "libsigk: " << sigk_last_error(h_) << "\n"; return rc; }
        input_on_device_ = true;
        return 0;
    }

    sigk_proteins packed() const {
        return sigk_proteins{residues_.data(), starts_.data(), func_.data(), seq_id_.data(), (uint64_t)func_.size()};
    }

    // tcc:183-213 on the GPU
    int ensure_handle(int device) {
        if (h_) return 0;
        sigk_config cfg{SIGK_ABI_VERSION, SIGK_K, device, 0, 1, 0};
        const int rc = sigk_create(&cfg, &h_);
        if (rc) std::cerr << "sigk_create: " << sigk_last_error(nullptr) << "\n";
        return rc;
    }
    int process_kmers(int device, sigk_table *table) {
        if (int rc = ensure_handle(device)) return rc;
        int rc = 0;
        if (!input_on_device_) {            // (extract_kmers_gpu has committed the proteins where they are)
            const sigk_proteins p = packed();
            rc = sigk_set_proteins(h_, &p);
        }
        if (!rc) rc = sigk_build(h_);
        if (!rc) rc = sigk_result(h_, table);
        if (rc) { std::cerr << "libsigk: " << sigk_last_error(h_) << "\n"; return rc; }
        std::cout << "Kept " << table->n_kept << " kmers\n";
        std::cout << "distinct_signatures=" << table->distinct_signatures << "\n";
        std::cout << "num_seqs_with_a_signature=" << table->num_seqs_with_a_signature << "\n";
        return 0;
    }
    // batch lookups against the table the build left on the device (the recall pass, function_caller.h)
    int lookup(const uint8_t *residues, const uint64_t *starts, uint64_t n_proteins, uint32_t *rows) {
        const int rc = sigk_lookup(h_, residues, starts, n_proteins, rows);
        if (rc) std::cerr << "libsigk: " << sigk_last_error(h_) << "\n";
        return rc;
    }
    ~HostSignatureBuilder() { if (h_) sigk_destroy(h_); }

    std::string lookup_function(uint16_t idx) const { return fm_.lookup_function(idx); }
    const std::vector<fs::path> &all_fasta_data() const { return all_fasta_data_; }
    const FunctionMap &function_map() const { return fm_; }

private:
    int n_threads_, max_seqs_per_file_;
    FunctionMap fm_;
    std::vector<fs::path> all_fasta_data_;
    // One sigk_fasta_parse over a set of files: the files' bytes (pinned, file i at begin[i], 16-byte aligned) and a
    // copy of the record table (the handle's arrays live only until the next parse).
    struct GpuFastaBatch {
        std::vector<fs::path> files;
        std::vector<uint64_t> begin, len;
        uint8_t *buf = nullptr;
        uint64_t n_records = 0, n_errors = 0;
        std::vector<uint64_t> header_pos, id_end, line_end, seq_begin, errors;
        std::vector<uint32_t> error_record;
        ~GpuFastaBatch() { if (buf) sigk_host_free(buf); }
        std::pair<uint64_t, uint64_t> records_of(size_t file) const {
            const auto lo = std::lower_bound(header_pos.begin(), header_pos.end(), begin[file]);
            const auto hi = std::lower_bound(header_pos.begin(), header_pos.end(), begin[file] + len[file]);
            return {(uint64_t)(lo - header_pos.begin()), (uint64_t)(hi - header_pos.begin())};
        }
        std::string text(uint64_t from, uint64_t to) const {                // bytes [from, to) minus '\r'
            std::string out;
            for (uint64_t p = from; p < to; ++p) if (buf[p] != '\r') out.push_back((char)buf[p]);
            return out;
        }
        std::string id_of(uint64_t r, size_t file) const {
            const uint64_t end = begin[file] + len[file];
            return text(header_pos[r] + 1, id_end[r] == SIGK_FASTA_NO_POS ? end : id_end[r]);
        }
        std::string def_of(uint64_t r, size_t file) const {
            const uint64_t end = begin[file] + len[file];
            if (id_end[r] == SIGK_FASTA_NO_POS) return std::string();
            return text(id_end[r], line_end[r] == SIGK_FASTA_NO_POS ? end : line_end[r]);
        }
    };
    int gpu_parse(const std::vector<fs::path> &files, GpuFastaBatch &b) {
        const size_t nf = files.size();
        b.files = files;
        b.begin.assign(nf, 0); b.len.assign(nf, 0);
        uint64_t at = 0;
        for (size_t i = 0; i < nf; ++i) {
            std::error_code ec;
            const auto size = fs::file_size(files[i], ec);
            b.begin[i] = at; b.len[i] = ec ? 0 : (uint64_t)size;
            at += (b.len[i] + 15) / 16 * 16;
        }
        b.buf = static_cast<uint8_t *>(sigk_host_alloc(at + 16));
        if (!b.buf) { std::cerr << "cannot allocate " << at << " bytes of pinned memory\n"; return SIGK_E_NOMEM; }
        parallel_for_index(nf, n_threads_, [&](size_t i) {
            std::ifstream in(files[i], std::ios::binary);
            in.read(reinterpret_cast<char *>(b.buf + b.begin[i]), (std::streamsize)b.len[i]);
            b.len[i] = (uint64_t)in.gcount();
        });
        sigk_fasta_records rec;
        const int rc = sigk_fasta_parse(h_, b.buf, b.begin.data(), b.len.data(), nf, &rec);
        if (rc) { std::cerr << "libsigk: " << sigk_last_error(h_) << "\n"; return rc; }
        b.n_records = rec.n_records; b.n_errors = rec.n_errors;
        b.header_pos.assign(rec.header_pos, rec.header_pos + rec.n_records);
        b.id_end.assign(rec.id_end, rec.id_end + rec.n_records);
        b.line_end.assign(rec.line_end, rec.line_end + rec.n_records);
        b.seq_begin.assign(rec.seq_begin, rec.seq_begin + rec.n_records + 1);
        const uint64_t ne = std::min<uint64_t>(rec.n_errors, SIGK_FASTA_MAX_ERRORS);
        b.errors.assign(rec.errors, rec.errors + ne);
        b.error_record.assign(rec.error_record, rec.error_record + ne);
        return 0;
    }
    // what the reference's parser prints for a character it drops (src/fasta_parser.h:135-138), file by file
    void report_fasta_errors(const GpuFastaBatch &b) const {
        size_t e = 0;
        for (size_t f = 0; f < b.files.size() && e < b.errors.size(); ++f) {
            const uint64_t end = b.begin[f] + b.len[f];
            uint64_t line = 1, counted = b.begin[f];
            for (; e < b.errors.size() && (b.errors[e] & ((1ull << 60) - 1)) < end; ++e) {
                const uint64_t pos = b.errors[e] & ((1ull << 60) - 1);
                const unsigned state = (unsigned)(b.errors[e] >> 60);
                for (; counted <= pos; ++counted) if (b.buf[counted] == '\n') ++line;      // the parser counts the newline before it looks at the state
                std::string what = state == 0 ? "Missing >" : (state == 3 ? "Bad data character '" : "Bad id or data character '");
                if (state != 0) { what += (char)b.buf[pos]; what += "'"; }
                const std::string id = (state == 0 || b.error_record[e] == 0xFFFFFFFFu) ? std::string() : b.id_of(b.error_record[e], f);
                std::cerr << "Error found: " << what << " at line " << line << " id='" << id << "'" << std::endl;
            }
        }
        if (b.n_errors > b.errors.size()) std::cerr << "(" << (b.n_errors - b.errors.size()) << " more parse errors not listed)\n";
    }
    std::unique_ptr<GpuFastaBatch> batch_;
    bool gpu_fasta_ = false;
    bool input_on_device_ = false;
    std::vector<uint8_t> residues_;
    std::vector<uint64_t> starts_;
    std::vector<uint16_t> func_;
    std::vector<uint32_t> seq_id_;
    sigk_handle *h_ = nullptr;
};

}  // namespace sigk_host
