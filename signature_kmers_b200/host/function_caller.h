// function_caller.h — the consumer side of the kept table: calling
// protein functions from signature k-mer hits, restated in plain C++17 (no Boost, no TBB).
//
//   SortedKmerDb     the KmerDb concept (src/call_functions.h:60-66; src/kept_kmer_db.h:10-32 for the in-memory
//                    map, src/cmph_kmer.h:27-147 for the on-disk perfect hash) over the table libsigk returns:
//                    rows come in the two sorted sections of include/sigk.h (k-mers without a lower-case residue in
//                    byte order, then the others by case-folded bytes and case mask), so a fetch is a binary search
//                    in the section the query's case pattern picks (bucketed in the first).  Also the
//                    kmer_data.sigk file that stands in for kmer_data.{mph,dat} (cmph is not in this image).
//   for_each_kmer    src/kmer_data.h:76-102, with its quirk: a window is skipped when it CONTAINS '*' / 'X'
//                    and also when it ENDS immediately before one (`kend >= next_ambig`).
//   FunctionCaller   src/call_functions.tcc:6-659 — HitSet::process (:34-101), process_aa_seq (:262-343),
//                    find_best_call (:352-659), process_fasta_stream (:220-258).
//
// Floating point: HitSet::process uses Boost.Math mean / median / median_absolute_deviation on float
// vectors (src/call_functions.tcc:49-51).  Boost is absent here, so those three are restated from the
// published algorithms (mean: four running means, Boost.Math >= 1.72 single_pass.hpp; median and MAD by
// selection, exact for any order).  Only the mean can differ from another Boost version, and only in the
// last ulp of a cut-off compared with the protein length — parity for that piece is UNPINNED against a
// stock reference run.  The logic itself is pinned: tests compare this header line by line with the
// reference's call_functions.tcc compiled from source over stand-in third-party headers, and with an
// independent Python restatement (tests/test_reference_shim.py, tests/test_function_caller.py).
#pragma once

#include "signature_host.h"

#include <algorithm>
#include <array>
#include <cstring>
#include <memory>
#include <sstream>

namespace sigk_host {

struct StoredKmerData {         // src/kmer_data.h:114-128
    uint16_t avg_from_end = 0;
    uint16_t function_index = 0xFFFF;
    uint16_t mean = 0;
    uint16_t median = 0;
    uint16_t var = 0;
};

constexpr uint16_t kUndefinedFunction = 0xFFFF;     // src/kmer_data.h:23
constexpr int kCallK = 8;

// ---------------------------------------------------------------------------
class SortedKmerDb {
public:
    static const int KmerSize = kCallK;

    SortedKmerDb() = default;
    // view over a table owned by somebody else (libsigk's result buffers)
    explicit SortedKmerDb(const sigk_table &t) { attach(t.n_kept, t.kmer, t.avg_from_end, t.function_index, t.mean, t.median, t.var); }

    uint64_t size() const { return n_; }
    const char *kmer_bytes() const { return kmer_; }

    // kmer_data.sigk: "SIGKTBL1", u64 rows, 8 rows bytes of k-mers, then the five u16 columns
    static bool write_file(const fs::path &file, const sigk_table &t) {
        std::FILE *f = std::fopen(file.c_str(), "wb");
        if (!f) return false;
        const uint64_t n = t.n_kept;
        bool ok = std::fwrite("SIGKTBL1", 1, 8, f) == 8 && std::fwrite(&n, 8, 1, f) == 1;
        ok = ok && std::fwrite(t.kmer, 8, n, f) == n;
        for (const uint16_t *col : {t.avg_from_end, t.function_index, t.mean, t.median, t.var}) ok = ok && std::fwrite(col, 2, n, f) == n;
        return std::fclose(f) == 0 && ok;
    }
    bool load_file(const fs::path &file) {
        std::FILE *f = std::fopen(file.c_str(), "rb");
        if (!f) return false;
        char magic[8];
        uint64_t n = 0;
        bool ok = std::fread(magic, 1, 8, f) == 8 && std::memcmp(magic, "SIGKTBL1", 8) == 0 && std::fread(&n, 8, 1, f) == 1;
        if (ok) {
            // the header's row count must agree with the file size (16 + 18 n) before anything is sized from it
            std::error_code fec;
            const uintmax_t sz = fs::file_size(file, fec);
            ok = !fec && n <= (UINT64_MAX - 16) / 18 && sz == 16 + 18 * (uintmax_t)n;
        }
        if (ok) {
            own_kmer_.resize(n * 8);
            for (auto &c : own_cols_) c.resize(n);
            ok = std::fread(own_kmer_.data(), 8, n, f) == n;
            for (auto &c : own_cols_) ok = ok && std::fread(c.data(), 2, n, f) == n;
        }
        std::fclose(f);
        if (ok) attach(n, own_kmer_.data(), own_cols_[0].data(), own_cols_[1].data(), own_cols_[2].data(), own_cols_[3].data(), own_cols_[4].data());
        return ok;
    }

    // row of the k-mer, or -1
    int64_t find(const char *kmer) const {
        if (!case_bits(kmer)) {
            // first section: byte order, bucketed by the leading two bytes
            const uint32_t b = bucket_of(kmer);
            uint64_t lo = bucket_[b], hi = bucket_[b + 1];
            while (lo < hi) {
                const uint64_t mid = lo + (hi - lo) / 2;
                const int c = std::memcmp(kmer_ + 8 * mid, kmer, 8);
                if (c == 0) return (int64_t)mid;
                if (c < 0) lo = mid + 1; else hi = mid;
            }
            return -1;
        }
        // second section: (case-folded bytes, case mask with residue j in bit j)
        const uint64_t qf = folded(kmer), qm = case_mask(kmer);
        uint64_t lo = n_upper_, hi = n_;
        while (lo < hi) {
            const uint64_t mid = lo + (hi - lo) / 2;
            const uint64_t rf = folded(kmer_ + 8 * mid), rm = case_mask(kmer_ + 8 * mid);
            if (rf == qf && rm == qm) return (int64_t)mid;
            if (rf < qf || (rf == qf && rm < qm)) lo = mid + 1; else hi = mid;
        }
        return -1;
    }
    uint64_t n_upper() const { return n_upper_; }
    StoredKmerData row(uint64_t i) const { return StoredKmerData{cols_[0][i], cols_[1][i], cols_[2][i], cols_[3][i], cols_[4][i]}; }

    template <class CB>
    void fetch(const std::array<char, kCallK> &k, CB cb, int &ec) const {
        const int64_t i = find(k.data());
        if (i >= 0) cb(row((uint64_t)i));
        ec = 0;
    }

private:
    uint64_t n_ = 0, n_upper_ = 0;          // rows; rows of the first section
    const char *kmer_ = nullptr;
    const uint16_t *cols_[5] = {};
    std::vector<uint64_t> bucket_;          // first row whose leading two bytes are >= the bucket's
    std::vector<char> own_kmer_;
    std::vector<uint16_t> own_cols_[5];

    static uint32_t bucket_of(const char *k) {
        return ((uint32_t)(unsigned char)k[0] << 8) | (uint32_t)(unsigned char)k[1];
    }
    static uint64_t load8(const char *k) { uint64_t v; std::memcpy(&v, k, 8); return v; }
    static uint64_t case_bits(const char *k) { return load8(k) & 0x2020202020202020ull; }
    static uint64_t folded(const char *k) {                 // bytes with the case bit cleared, first residue most significant
        return __builtin_bswap64(load8(k) & ~0x2020202020202020ull);
    }
    static uint64_t case_mask(const char *k) {              // bit j set iff residue j is lower case
        uint64_t m = 0;
        for (int j = 0; j < 8; ++j) m |= (uint64_t)(((unsigned char)k[j] >> 5) & 1u) << j;
        return m;
    }
    void attach(uint64_t n, const char *kmer, const uint16_t *a, const uint16_t *f, const uint16_t *m, const uint16_t *md, const uint16_t *v) {
        n_ = n; kmer_ = kmer;
        cols_[0] = a; cols_[1] = f; cols_[2] = m; cols_[3] = md; cols_[4] = v;
        // the second section starts at the first row with a lower-case residue (a monotone predicate over the rows)
        uint64_t lo = 0, hi = n;
        while (lo < hi) {
            const uint64_t mid = lo + (hi - lo) / 2;
            if (case_bits(kmer + 8 * mid)) hi = mid; else lo = mid + 1;
        }
        n_upper_ = lo;
        bucket_.assign((1u << 16) + 1, 0);
        // the first section is sorted by bytes: count per leading 2-byte prefix, then prefix sums
        for (uint64_t i = 0; i < n_upper_; ++i) ++bucket_[bucket_of(kmer + 8 * i) + 1];
        for (size_t b = 1; b < bucket_.size(); ++b) bucket_[b] += bucket_[b - 1];
    }
};

// ---------------------------------------------------------------------------
// src/kmer_data.h:76-102
template <class F>
void for_each_kmer(const std::string &s, F cb) {
    const size_t n = s.size();
    if (n < (size_t)kCallK) return;
    auto next_ambig_from = [&](size_t from) {
        for (size_t i = from; i < n; ++i)
            if (s[i] == '*' || s[i] == 'X') return i;
        return n;
    };
    size_t amb = next_ambig_from(0);
    size_t p = 0;
    std::array<char, kCallK> kmer;
    while (p + kCallK <= n) {
        const size_t kend = p + kCallK;
        if (amb != n && kend >= amb) {            // the window touches, or ends right before, the ambiguity
            p = amb + 1;
            amb = next_ambig_from(p);
            continue;
        }
        std::memcpy(kmer.data(), s.data() + p, kCallK);
        cb(kmer, p);
        ++p;
    }
}

// ---------------------------------------------------------------------------
// Boost.Math univariate statistics on floats, as used at src/call_functions.tcc:49-51
namespace callstats {

inline float mean(const std::vector<float> &v) {
    const size_t n = v.size();
    float mu[4] = {0.f, 0.f, 0.f, 0.f};
    float i = 1.f;
    const size_t body = n - (n % 4);
    size_t k = 0;
    for (; k < body; k += 4) {
        const float inv = 1.f / i;
        for (int j = 0; j < 4; ++j) { float t = v[k + j] - mu[j]; t *= inv; mu[j] += t; }
        i += 1.f;
    }
    const float num1 = (float)body / 4.f;
    const float num2 = num1 + (float)(n % 4);
    for (; k < n; ++k) { mu[3] += (v[k] - mu[3]) / i; i += 1.f; }
    return (num1 * ((mu[0] + mu[1]) + mu[2]) + num2 * mu[3]) / (float)n;
}

inline float median(std::vector<float> &v) {        // permutes v, like the original
    const size_t n = v.size();
    if (n % 2 == 0) {
        auto mid = v.begin() + (n / 2 - 1);
        std::nth_element(v.begin(), mid, v.end());
        const float hi = *std::min_element(mid + 1, v.end());
        return (*mid + hi) / 2.f;
    }
    auto mid = v.begin() + n / 2;
    std::nth_element(v.begin(), mid, v.end());
    return *mid;
}

inline float median_absolute_deviation(std::vector<float> &v) {
    const float center = median(v);
    const size_t n = v.size();
    auto cmp = [center](float a, float b) { return std::fabs(a - center) < std::fabs(b - center); };
    if (n % 2 == 0) {
        auto mid = v.begin() + (n / 2 - 1);
        std::nth_element(v.begin(), mid, v.end(), cmp);
        const float hi = *std::min_element(mid + 1, v.end(), cmp);
        return (std::fabs(*mid - center) + std::fabs(hi - center)) / 2.f;
    }
    auto mid = v.begin() + n / 2;
    std::nth_element(v.begin(), mid, v.end(), cmp);
    return std::fabs(*mid - center);
}

}  // namespace callstats

// ---------------------------------------------------------------------------
struct KmerCall {                   // src/call_functions.h:23-48
    unsigned start = 0, end = 0;
    int count = 0;
    uint16_t function_index = kUndefinedFunction;
    unsigned protein_length_median = 0;
    float protein_length_med_avg_dev = 0.f;
};

struct BestCall {
    uint16_t function_index = kUndefinedFunction;
    std::string function;
    float score = 0.f;
    float score_offset = 0.f;
};

template <class KmerDb>
class FunctionCaller {
public:
    FunctionCaller(const KmerDb &db, const fs::path &function_index_file, int min_hits = 5, int max_gap = 200)
        : db_(db), min_hits_(min_hits), max_gap_(max_gap) {
        read_function_index(function_index_file);
    }
    FunctionCaller(const KmerDb &db, std::vector<std::string> function_index, int min_hits = 5, int max_gap = 200)
        : db_(db), min_hits_(min_hits), max_gap_(max_gap), function_index_(std::move(function_index)) { find_hypothetical(); }

    void ignore_hypothetical(bool x) { ignore_hypothetical_ = x; }
    const std::vector<std::string> &function_index() const { return function_index_; }
    const std::string &function_at_index(int idx) const {
        return idx == kUndefinedFunction ? undefined_function_ : function_index_[(size_t)idx];
    }

    // src/call_functions.tcc:123-147: two reads of the file, ids may come in any order
    void read_function_index(const fs::path &file) {
        function_index_.clear();
        std::ifstream in(file);
        std::string line;
        std::vector<std::pair<int, std::string>> rows;
        int max_id = 0;
        while (std::getline(in, line)) {
            const size_t tab = line.find('\t');
            const int id = std::stoi(line.substr(0, tab));
            max_id = std::max(max_id, id);
            std::string name;
            if (tab != std::string::npos) {
                const size_t tab2 = line.find('\t', tab + 1);
                name = line.substr(tab + 1, tab2 == std::string::npos ? std::string::npos : tab2 - tab - 1);
            }
            rows.emplace_back(id, name);
        }
        function_index_.resize((size_t)max_id + 1);
        for (auto &r : rows) function_index_[(size_t)r.first] = r.second;
        find_hypothetical();
    }

    // src/call_functions.tcc:262-343.  hit_cb(id, kmer, offset, seqlen, kdata) sees every accepted hit.
    template <class HitCB>
    void process_aa_seq(const std::string &id, const std::string &seq, std::vector<KmerCall> *calls, HitCB hit_cb) const {
        HitRun run(*this, id, (double)seq.size(), calls);
        for_each_kmer(seq, [&](const std::array<char, kCallK> &kmer, size_t offset) {
            int ec = 0;
            db_.fetch(kmer, [&](const StoredKmerData &kd) { run.add(kmer, offset, kd, hit_cb); }, ec);
        });
        run.finish();
    }

    // The same with the lookups already done in a batch (libsigk's sigk_lookup): rows[p] = table row of the window
    // at position p of this protein, or 0xFFFFFFFF when the window is not visited or not in the table.
    template <class HitCB>
    void process_looked_up(const std::string &id, const std::string &seq, const uint32_t *rows, std::vector<KmerCall> *calls,
                           HitCB hit_cb) const {
        HitRun run(*this, id, (double)seq.size(), calls);
        std::array<char, kCallK> kmer;
        for (size_t p = 0; p + kCallK <= seq.size(); ++p) {
            if (rows[p] == 0xFFFFFFFFu) continue;
            std::memcpy(kmer.data(), seq.data() + p, kCallK);
            run.add(kmer, p, db_.row(rows[p]), hit_cb);
        }
        run.finish();
    }

    // src/call_functions.tcc:352-659
    BestCall find_best_call(const std::vector<KmerCall> &calls) const {
        BestCall out;
        if (calls.empty()) return out;

        // adjacent calls with the same function become one
        std::vector<KmerCall> collapsed;
        for (size_t i = 0; i < calls.size();) {
            KmerCall cur = calls[i++];
            while (i < calls.size() && calls[i].function_index == cur.function_index) {
                cur.end = calls[i].end;
                cur.count += calls[i].count;
                ++i;
            }
            collapsed.push_back(cur);
        }
        // F1 | weak F2 | F1  ->  one F1 call (interior below 5 hits, the two outer ones 10 or more together)
        std::vector<KmerCall> merged;
        for (size_t i = 0; i < collapsed.size();) {
            KmerCall cur = collapsed[i++];
            while (i < collapsed.size() && i + 1 < collapsed.size() && cur.function_index == collapsed[i + 1].function_index &&
                   collapsed[i].count < 5 && cur.count + collapsed[i + 1].count >= 10) {
                cur.end = collapsed[i + 1].end;
                cur.count += collapsed[i + 1].count;
                i += 2;
            }
            merged.push_back(cur);
        }

        if (merged.size() > 1 && try_fusion(merged, out)) return out;

        // total hits per function, the two largest decide
        std::map<int, int> by_func;
        for (const auto &c : merged) by_func[c.function_index] += c.count;
        std::vector<std::pair<int, int>> vec(by_func.begin(), by_func.end());
        // Exactly the reference's call (:594-599): only the first two places are sorted, and the code below reads
        // vec[2] all the same — which element std::partial_sort leaves there is unspecified by the standard but
        // fixed for a given standard library, and it decides whether "f1 ?? f2" is named when the two best are
        // within 5 hits.  The same call on the same input order (the std::map walk above) reproduces it under
        // libstdc++, the reference's library (g++, Makefile:42); found by fuzzing against the reference's sources.
        if (vec.size() > 1)
            std::partial_sort(vec.begin(), vec.begin() + 2, vec.end(), [](const std::pair<int, int> &a, const std::pair<int, int> &b) { return a.second > b.second; });

        out.score_offset = vec.size() == 1 ? (float)vec[0].second : (float)(vec[0].second - vec[1].second);
        if (out.score_offset >= 5.0f) {
            out.function_index = (uint16_t)vec[0].first;
            out.function = function_at_index(vec[0].first);
            out.score = (float)vec[0].second;
            return out;
        }
        // too close to call: name the two best if they stand clear of the third
        if (vec.size() >= 2) {
            std::string f1 = function_at_index(vec[0].first), f2 = function_at_index(vec[1].first);
            if (f2 > f1) std::swap(f1, f2);
            if (vec.size() == 2) {
                out.function = f1 + " ?? " + f2;
                out.score = (float)vec[0].second;
            } else {
                const float pair_offset = (float)(vec[1].second - vec[2].second);
                if (pair_offset > 2.0f) {
                    out.function = f1 + " ?? " + f2;
                    out.score = (float)vec[0].second;
                    out.score_offset = pair_offset;
                }
            }
        }
        return out;
    }

    // src/call_functions.tcc:220-258: every FASTA record -> hits -> calls -> best call -> call_cb
    template <class HitCB, class CallCB>
    void process_fasta_stream(std::istream &in, HitCB &hit_cb, CallCB &call_cb) const {
        FastaReader reader([&](const std::string &id, const std::string &, const std::string &seq) {
            if (id.empty()) return;
            std::vector<KmerCall> calls;
            process_aa_seq(id, seq, &calls, hit_cb);
            const BestCall best = find_best_call(calls);
            call_cb(id, best.function, best.function_index, best.score, seq.size());
        }, true);
        reader.parse(in);
        reader.finish();
    }

    // process_fasta_stream with the lookups of the whole stream done in one batch: lookup(residues, starts,
    // n_proteins, rows) fills rows as libsigk's sigk_lookup does and returns 0 on success.  Returns its status.
    template <class Lookup, class HitCB, class CallCB>
    int process_fasta_stream_batched(std::istream &in, Lookup &&lookup, HitCB &hit_cb, CallCB &call_cb) const {
        std::vector<std::string> ids, seqs;
        FastaReader reader([&](const std::string &id, const std::string &, const std::string &seq) {
            if (id.empty()) return;
            ids.push_back(id);
            seqs.push_back(seq);
        }, true);
        reader.parse(in);
        reader.finish();
        std::vector<uint8_t> residues;
        std::vector<uint64_t> starts(1, 0);
        for (const auto &s : seqs) { residues.insert(residues.end(), s.begin(), s.end()); starts.push_back(residues.size()); }
        std::vector<uint32_t> rows(residues.size() + 1, 0xFFFFFFFFu);
        if (const int rc = lookup(residues.data(), starts.data(), (uint64_t)seqs.size(), rows.data())) return rc;
        for (size_t i = 0; i < seqs.size(); ++i) {
            std::vector<KmerCall> calls;
            process_looked_up(ids[i], seqs[i], rows.data() + starts[i], &calls, hit_cb);
            const BestCall best = find_best_call(calls);
            call_cb(ids[i], best.function, best.function_index, best.score, seqs[i].size());
        }
        return 0;
    }

private:
    struct Hit { StoredKmerData kd; unsigned long pos; };

    const KmerDb &db_;
    int min_hits_, max_gap_;
    bool ignore_hypothetical_ = false;
    std::vector<std::string> function_index_;
    std::string undefined_function_;
    int hypo_pos_ = -1;

    // the per-hit state machine of process_aa_seq (:276-339), shared by both hit sources
    class HitRun {
    public:
        HitRun(const FunctionCaller &fc, const std::string &id, double seqlen, std::vector<KmerCall> *calls)
            : fc_(fc), id_(id), seqlen_(seqlen), calls_(calls) {
            if (fc.hypo_pos_ < 0) {
                std::cerr << "Cannot find hypothetical protein index\n";
                std::exit(1);
            }
        }
        template <class HitCB>
        void add(const std::array<char, kCallK> &kmer, size_t offset, const StoredKmerData &kd, HitCB &hit_cb) {
            if (fc_.ignore_hypothetical_ && kd.function_index == fc_.hypo_pos_) return;
            hit_cb(id_, kmer, offset, seqlen_, kd);
            // a hit further than max_gap from the last one closes the current run of hits
            if (!hits_.empty() && hits_.back().pos + (unsigned long)fc_.max_gap_ < offset) {
                if ((int)hits_.size() >= fc_.min_hits_) fc_.process_hits(hits_, seqlen_, current_, calls_);
                else hits_.clear();
            }
            if (hits_.empty()) current_ = kd.function_index;
            // order_constraint_ is hard-wired false (:115), so every hit is taken
            hits_.push_back(Hit{kd, (unsigned long)offset});
            // two hits in a row for another function: close the run, start the next with those two
            if (hits_.size() > 1 && current_ != kd.function_index) {
                const size_t m = hits_.size();
                if (hits_[m - 2].kd.function_index == hits_[m - 1].kd.function_index) fc_.process_hits(hits_, seqlen_, current_, calls_);
            }
        }
        void finish() {
            if ((int)hits_.size() >= fc_.min_hits_) fc_.process_hits(hits_, seqlen_, current_, calls_);
        }

    private:
        const FunctionCaller &fc_;
        const std::string &id_;
        double seqlen_;
        std::vector<KmerCall> *calls_;
        std::vector<Hit> hits_;
        uint16_t current_ = kUndefinedFunction;
    };

    void find_hypothetical() {
        hypo_pos_ = -1;
        for (size_t i = 0; i < function_index_.size(); ++i)
            if (function_index_[i] == "hypothetical protein") { hypo_pos_ = (int)i; break; }
    }

    // HitSet::process, src/call_functions.tcc:34-101
    void process_hits(std::vector<Hit> &hits, double seqlen, uint16_t &current, std::vector<KmerCall> *calls) const {
        int count = 0;
        size_t last = 0;
        std::vector<float> lengths;
        for (size_t i = 0; i < hits.size(); ++i) {
            if (hits[i].kd.function_index == current) {
                last = i;
                ++count;
                lengths.push_back((float)hits[i].kd.mean);
            }
        }
        const float mean_length = callstats::mean(lengths);
        const float median_length = callstats::median(lengths);
        float mad_length = callstats::median_absolute_deviation(lengths);
        if (mad_length == 0) mad_length = 30;
        const double cutoff_b = mean_length - 2.0 * mad_length, cutoff_t = mean_length + 2.0 * mad_length;
        if (count >= min_hits_ && !(seqlen < cutoff_b || seqlen > cutoff_t) && calls)
            calls->push_back(KmerCall{(unsigned)hits[0].pos, (unsigned)(hits[last].pos + (kCallK - 1)), count, current,
                                      (unsigned)median_length, mad_length});
        const size_t m = hits.size();
        if (hits[m - 2].kd.function_index != current && hits[m - 2].kd.function_index == hits[m - 1].kd.function_index) {
            current = hits[m - 2].kd.function_index;
            hits.erase(hits.begin(), hits.end() - 2);
        } else {
            hits.clear();
        }
    }

    // The fusion test, src/call_functions.tcc:458-556: functions get letters in order of appearance ('A'..),
    // "X / Y" compound functions get fusion letters ('W'..); the letter string must match
    // ^W?A[A|W]*W[B|W]*BW?$ and the part lengths must add up: |mean(A) + mean(B) - mean(W)| / mean(W) < 0.1.
    bool try_fusion(const std::vector<KmerCall> &merged, BestCall &out) const {
        char next_func_key = 'A', next_fusion_key = 'W';
        std::map<std::string, char> func_map, fusion_map;
        std::map<char, std::pair<uint16_t, std::string>> key_info;
        struct Acc { float sum = 0.f; size_t n = 0; };
        std::map<char, Acc> part;
        std::string exp;
        int sum_scores = 0;
        for (const auto &c : merged) {
            sum_scores += c.count;
            const std::string func = function_at_index(c.function_index);
            std::vector<std::string> parts;
            for (size_t start = 0;;) {          // operators.h:80-91, delimiter " / "
                const size_t e = func.find(" / ", start);
                parts.push_back(func.substr(start, e == std::string::npos ? std::string::npos : e - start));
                if (e == std::string::npos) break;
                start = e + 3;
            }
            std::string fusion_key;
            for (const auto &p : parts) {
                if (!func_map.count(p)) func_map[p] = next_func_key++;
                fusion_key += func_map[p];
            }
            char key;
            if (parts.size() > 1) {
                if (!fusion_map.count(fusion_key)) fusion_map[fusion_key] = next_fusion_key++;
                key = fusion_map[fusion_key];
            } else {
                key = func_map[func];
            }
            exp += key;
            Acc &a = part[key];
            a.sum += (float)c.protein_length_median;
            ++a.n;
            key_info[key] = {c.function_index, func};
        }
        if (!fusion_pattern(exp)) return false;
        auto mean_of = [&](char k) { const Acc &a = part[k]; return a.sum / (float)a.n; };      // accumulators mean: sum / count
        const float a_mean = mean_of('A'), w_mean = mean_of('W'), b_mean = mean_of('B');
        const float diff = (a_mean + b_mean) - w_mean;
        // `abs(diff)` in the reference is unqualified; with <cmath>'s overloads visible it is the float one
        const float frac = std::fabs(diff) / w_mean;
        if (!(frac < 0.1)) return false;
        out.function_index = key_info['W'].first;
        out.function = key_info['W'].second;
        out.score = (float)sum_scores;
        out.score_offset = 0.f;
        return true;
    }

    // full match of ^W?A[A|W]*W[B|W]*BW?$ ('|' is a literal member of both classes)
    static bool fusion_pattern(const std::string &s) {
        const size_t n = s.size();
        for (size_t p0 = 0; p0 <= 1 && p0 <= n; ++p0) {
            if (p0 == 1 && s[0] != 'W') break;
            size_t i = p0;
            if (i >= n || s[i] != 'A') continue;
            ++i;
            // [A|W]* then a 'W': try every 'W' reachable through [A|W]*
            for (size_t w = i; w < n && (s[w] == 'A' || s[w] == '|' || s[w] == 'W'); ++w) {
                if (s[w] != 'W') continue;
                // [B|W]* then 'B' then W?$
                for (size_t b = w + 1; b < n && (s[b] == 'B' || s[b] == '|' || s[b] == 'W'); ++b) {
                    if (s[b] != 'B') continue;
                    if (b + 1 == n || (b + 2 == n && s[b + 1] == 'W')) return true;
                }
            }
        }
        return false;
    }
};

// ---------------------------------------------------------------------------
// Output files of kmers-build-signatures that are derived from the finished table.

// distinct_functions, src/kmers-build-signatures.cc:230-236: idx \t function \t kept k-mers (here in index order; the
// reference walks a hash map)
inline void write_distinct_functions(const fs::path &file, const sigk_table &t, const FunctionMap &fm) {
    std::ofstream df(file);
    for (unsigned f = 0; f < 65536; ++f)
        if (t.distinct_functions[f]) df << f << "\t" << fm.lookup_function((uint16_t)f) << "\t" << t.distinct_functions[f] << "\n";
}

// recall.report.d/<fasta file>, src/kmers-build-signatures.cc:266-349: the source proteins are called with the new
// k-mers; every call that differs from the stripped original assignment is reported, ordered by id:
// id \t original \t original stripped \t call \t function_index \t score.  `lookup` (optional) does the k-mer lookups of a
// whole file in one batch (libsigk's sigk_lookup); without it every window is fetched from the table on the host.
using BatchLookup = std::function<int(const uint8_t *, const uint64_t *, uint64_t, uint32_t *)>;
inline bool write_recall_reports(const FunctionMap &fm, const std::vector<fs::path> &files, const sigk_table &t,
                                 const fs::path &function_index_file, const fs::path &report_dir, int n_threads, const BatchLookup &lookup) {
    const SortedKmerDb kdb(t);
    const FunctionCaller<SortedKmerDb> caller(kdb, function_index_file);
    std::atomic<bool> failed{false};
    parallel_for_index(files.size(), n_threads, [&](size_t i) {
        struct Row { std::string old_func, old_stripped, new_func; int func_index; float score; };
        std::map<std::string, Row> rows;
        auto hit_cb = [](const std::string &, const std::array<char, kCallK> &, size_t, double, const StoredKmerData &) {};
        auto call_cb = [&](const std::string &id, const std::string &func, uint16_t fi, float score, size_t) {
            std::string orig, orig_stripped;
            fm.lookup_original_assignment(id, orig, orig_stripped);
            if (orig_stripped != func) rows.emplace(id, Row{orig, orig_stripped, func, (int)fi, score});
        };
        std::ifstream in(files[i]);
        if (lookup) { if (caller.process_fasta_stream_batched(in, lookup, hit_cb, call_cb)) failed = true; }
        else caller.process_fasta_stream(in, hit_cb, call_cb);
        std::ofstream rep(report_dir / files[i].filename());
        for (const auto &e : rows)
            rep << e.first << "\t" << e.second.old_func << "\t" << e.second.old_stripped << "\t" << e.second.new_func << "\t"
                << e.second.func_index << "\t" << e.second.score << "\n";
    });
    return !failed;
}

}  // namespace sigk_host
