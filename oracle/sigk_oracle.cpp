/*
 * sigk_oracle.cpp — CPU restatement of the reference's signature-generation
 * hot path.  TEST INFRASTRUCTURE ONLY.  Pinned against the reference's own sources compiled over
 * stand-in third-party headers (oracle/ref_signature_shim.cpp, tests/test_reference_shim.py);
 * PARITY UNPINNED against a stock reference run for the Boost/TBB internals (see sigk_oracle.h).
 *
 * Follows, function by function (paths relative to the reference checkout):
 *   extract()            src/signature_build.tcc:47-70   (extract_kmers)
 *   load_sequence()      src/signature_build.tcc:120-181 (load_kmers_from_sequence)
 *   process()            src/signature_build.tcc:183-213 (process_kmers)
 *   process_kmer_set()   src/signature_build.tcc:218-293
 *   BoostAcc             Boost.Accumulators sum.hpp / mean.hpp / median.hpp +
 *                        p_square_quantile.hpp / variance.hpp (published
 *                        algorithms; Boost is not in this image)
 *   Table                tbb::concurrent_unordered_multimap as used at
 *                        src/signature_build.h:62, :120 (equal keys iterate
 *                        newest-first; TBB <= 2020 internal_insert)
 *
 * The data structure deliberately keeps the reference's shape (one heap node
 * per k-mer occurrence in a chained hash table, then a per-key walk) because
 * bench.py times this code as the CPU baseline.
 *
 * Build: g++ -O3 -g -std=c++17 -ffp-contract=off -fPIC -shared (oracle/Makefile);
 * -O3 -g are the reference's flags (Makefile:42-48); no -march, so no FMA.
 */
#include "sigk_oracle.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <thread>
#include <vector>

namespace {

/* src/kmer_data.h:105-112 */
struct KmerAttributes {
    uint16_t func_index;
    uint16_t otu_index;
    unsigned short offset;
    unsigned int seq_id;
    unsigned int protein_length;
};
static_assert(sizeof(KmerAttributes) == 16, "layout of src/kmer_data.h:105-112");

struct Node {
    uint64_t kmer;          /* 8 raw chars, first char in the top byte: integer order == byte order */
    KmerAttributes attr;
    uint32_t next;          /* next node in the bucket chain (older) */
};

constexpr uint32_t NIL = 0xFFFFFFFFu;

struct Row {
    uint64_t kmer;
    uint16_t avg_from_end, function_index, mean, median, var;
};

/* src/signature_build.h:102-103 (ok_prot_) */
struct OkProt {
    bool ok[256];
    OkProt() {
        std::memset(ok, 0, sizeof ok);
        const char *aa = "ACDEFGHIKLMNPQRSTVWYacdefghiklmnpqrstvwy";
        for (const char *c = aa; *c; ++c) ok[(unsigned char)*c] = true;
    }
};
const OkProt g_ok;

inline uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

/* (unsigned short)double on x86-64, g++ -O3: cvttsd2si r32 then truncate.
 * Out-of-range / NaN produce the "integer indefinite" 0x80000000 -> 0.      */
inline uint16_t u16_from_double(double d) {
    int32_t i;
    if (!(d > -2147483649.0 && d < 2147483648.0)) i = INT32_MIN;   /* also catches NaN */
    else i = (int32_t)d;                                           /* truncation toward zero */
    return (uint16_t)(uint32_t)i;
}

/* accumulator_set<unsigned short, stats<tag::mean, tag::median, tag::variance>>
 * (src/signature_build.tcc:262-264).  Sample type unsigned short, samples
 * arrive as unsigned int (acc(item.protein_length), :271).                  */
struct BoostAcc {
    size_t n = 0;
    unsigned short S = 0;           /* sum_impl<unsigned short>: wraps mod 65536 */
    double imm_mean = 0.0;          /* only for the IMMEDIATE_MEAN reading */
    double var = 0.0;
    double q[5] = {0, 0, 0, 0, 0};
    double pos[5] = {1, 2, 3, 4, 5};
    double des[5] = {1, 2, 3, 4, 5};    /* 1, 1+2p, 1+4p, 3+2p, 5 with p = .5 */
    bool immediate = false;

    void push(unsigned int x) {
        static const double inc[5] = {0.0, 0.25, 0.5, 0.75, 1.0};  /* 0, p/2, p, (1+p)/2, 1 */
        n += 1;
        S = (unsigned short)(S + x);
        if (immediate) imm_mean = (imm_mean * (double)(n - 1) + (double)x) / (double)n;

        /* p_square_quantile_impl<unsigned short, for_median>::operator() */
        if (n <= 5) {
            q[n - 1] = (double)x;
            if (n == 5) std::sort(q, q + 5);
        } else {
            const double xd = (double)x;
            size_t k;
            if (xd < q[0]) { q[0] = xd; k = 1; }
            else if (q[4] <= xd) { q[4] = xd; k = 4; }
            else k = (size_t)(std::upper_bound(q, q + 5, xd) - q);
            for (size_t i = k; i < 5; ++i) pos[i] += 1.0;
            for (size_t i = 0; i < 5; ++i) des[i] += inc[i];
            for (size_t i = 1; i <= 3; ++i) {
                const double d = des[i] - pos[i];
                const double dp = pos[i + 1] - pos[i];
                const double dm = pos[i - 1] - pos[i];
                const double hp = (q[i + 1] - q[i]) / dp;
                const double hm = (q[i - 1] - q[i]) / dm;
                if ((d >= 1.0 && dp > 1.0) || (d <= -1.0 && dm < -1.0)) {
                    const short sign_d = (short)(d / std::fabs(d));
                    const double s = (double)sign_d;
                    /* heights[i] + sign_d / (dp - dm) * ((sign_d - dm) * hp + (dp - sign_d) * hm) */
                    const double a = s / (dp - dm);
                    const double t1 = (s - dm) * hp;
                    const double t2 = (dp - s) * hm;
                    const double h = q[i] + a * (t1 + t2);
                    if (q[i - 1] < h && h < q[i + 1]) q[i] = h;
                    else {
                        if (d > 0) q[i] += hp;
                        if (d < 0) q[i] -= hm;
                    }
                    pos[i] += s;
                }
            }
        }

        /* variance_impl<unsigned short, tag::mean, tag::sample>::operator() */
        if (n > 1) {
            const double mean_n = immediate ? imm_mean : (double)S / (double)n;
            const double tmp = (double)x - mean_n;
            var = (var * (double)(n - 1)) / (double)n + (tmp * tmp) / (double)(n - 1);
        }
    }
    double mean_f64() const { return immediate ? imm_mean : (double)S / (double)n; }
    double median_f64() const { return q[2]; }
};

struct Stats {
    uint64_t n_distinct = 0;
    uint64_t distinct_signatures = 0;
    std::vector<uint32_t> distinct_functions = std::vector<uint32_t>(SIGK_N_FUNCTION_SLOTS, 0);
};

}  // namespace

struct sigk_oracle {
    std::vector<Row> rows;
    std::vector<char> kmer_bytes;
    std::vector<uint16_t> avg, func, mean, median, var;
    std::vector<uint32_t> distinct_functions, seqs_with_func;
    uint64_t n_occurrences = 0, n_distinct = 0, distinct_signatures = 0, n_seqs_sig = 0;
    double t_extract = 0, t_process = 0;
};

namespace {

struct Table {
    std::vector<Node> nodes;
    std::vector<std::atomic<uint32_t>> heads;
    uint64_t mask = 0;
};

/* src/signature_build.tcc:162-180: the window loop of load_kmers_from_sequence.
 * Returns the number of records inserted.                                    */
uint64_t load_sequence(Table &t, uint64_t node_base, const uint8_t *seq, uint64_t len,
                       uint16_t function_index, uint32_t seq_id) {
    uint64_t n_ins = 0;
    if (len < SIGK_K) return 0;                     /* it < seq.end()-K+1 never true */
    for (uint64_t p = 0; p + SIGK_K <= len; ++p) {
        const unsigned short n = (unsigned short)(len - p);     /* :164 */
        bool ok = true;
        uint64_t kmer = 0;
        for (int j = 0; j < SIGK_K; ++j) {
            const unsigned char c = seq[p + j];
            if (!g_ok.ok[c]) { ok = false; break; }             /* :168-175 */
            kmer = (kmer << 8) | c;
        }
        if (!ok) continue;
        /* :178  insert({kmer, {function_index, UndefinedOTU, n, seq_id, seq.length()}}) */
        const uint64_t idx = node_base + n_ins++;
        Node &nd = t.nodes[idx];
        nd.kmer = kmer;
        nd.attr = KmerAttributes{function_index, 0xFFFF, n, seq_id, (unsigned int)len};
        /* newest-first inside a key: push at the head of the bucket chain */
        nd.next = t.heads[mix64(kmer) & t.mask].exchange((uint32_t)idx, std::memory_order_relaxed);
    }
    return n_ins;
}

/* src/signature_build.tcc:218-293 */
void process_kmer_set(uint64_t kmer, const Node *const *items, size_t count, bool immediate,
                      std::vector<Row> &kept, Stats &st, std::vector<std::atomic<uint64_t>> &seq_bits,
                      std::vector<unsigned short> &offsets, std::map<uint16_t, int> &func_count) {
    func_count.clear();
    for (size_t i = 0; i < count; ++i) func_count[items[i]->attr.func_index]++;    /* :203 */

    uint16_t best_func_1 = 0xFFFF, best_func_2 = 0xFFFF;
    int best_count_1 = -1, best_count_2 = -1;
    for (const auto &x : func_count) {                  /* :228-248, ascending index, strict > */
        if (best_func_1 == 0xFFFF) { best_func_1 = x.first; best_count_1 = x.second; }
        else if (x.second > best_count_1) {
            best_func_2 = best_func_1; best_count_2 = best_count_1;
            best_func_1 = x.first; best_count_1 = x.second;
        } else if (x.second > best_count_2) { best_func_2 = x.first; best_count_2 = x.second; }
    }
    (void)best_func_2;

    const float thresh = float((int)count) * 0.8f;      /* :250 */
    const int best_count = best_count_1;
    const uint16_t best_func = best_func_1;
    if ((float)best_count < thresh) return;             /* :254-257 */

    offsets.clear();
    BoostAcc acc;
    acc.immediate = immediate;
    for (size_t i = 0; i < count; ++i) {                /* :266-275, multimap iteration order */
        const KmerAttributes &item = items[i]->attr;
        if (item.func_index == best_func) acc.push(item.protein_length);
        offsets.push_back(item.offset);
        seq_bits[item.seq_id >> 6].fetch_or(1ULL << (item.seq_id & 63), std::memory_order_relaxed);
    }
    const uint16_t mean = u16_from_double(acc.mean_f64());      /* :277 */
    const uint16_t median = u16_from_double(acc.median_f64());  /* :278 */
    const uint16_t var = u16_from_double(acc.var);              /* :279 */

    std::sort(offsets.begin(), offsets.end());                  /* :281 */
    const unsigned short avg_from_end = offsets[offsets.size() / 2];

    st.distinct_signatures++;                                   /* :285 */
    st.distinct_functions[best_func]++;                         /* :286 */
    kept.push_back(Row{kmer, avg_from_end, best_func, mean, median, var});   /* :288 */
}

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" {

int sigk_oracle_build(const sigk_proteins *p, int n_threads, int flags, sigk_oracle **out) {
    if (!p || !out) return SIGK_E_INVALID;
    if (n_threads < 1) n_threads = 1;
    const uint64_t np = p->n_proteins;
    const bool immediate = (flags & SIGK_ORACLE_IMMEDIATE_MEAN) != 0;

    sigk_oracle *o = new (std::nothrow) sigk_oracle;
    if (!o) return SIGK_E_NOMEM;
    o->seqs_with_func.assign(SIGK_N_FUNCTION_SLOTS, 0);
    o->distinct_functions.assign(SIGK_N_FUNCTION_SLOTS, 0);

    /* node slots: protein i owns [base[i], base[i] + max(0, len-7)) */
    std::vector<uint64_t> base(np + 1, 0);
    uint32_t max_seq_id = 0;
    for (uint64_t i = 0; i < np; ++i) {
        const uint64_t len = p->starts[i + 1] - p->starts[i];
        base[i + 1] = base[i] + (len >= SIGK_K ? len - SIGK_K + 1 : 0);
        max_seq_id = std::max(max_seq_id, p->seq_id[i]);
        if (p->function_index[i] == SIGK_UNDEFINED_FUNCTION) { delete o; return SIGK_E_INVALID; }
    }
    const uint64_t cap = base[np];
    if (cap >= NIL) { delete o; return SIGK_E_UNSUPPORTED; }

    Table t;
    try {
        t.nodes.resize(cap);
        uint64_t nb = 1024;
        while (nb < cap) nb <<= 1;
        t.mask = nb - 1;
        t.heads = std::vector<std::atomic<uint32_t>>(nb);
    } catch (const std::bad_alloc &) { delete o; return SIGK_E_NOMEM; }
    for (auto &h : t.heads) h.store(NIL, std::memory_order_relaxed);

    std::vector<std::atomic<uint64_t>> seq_bits(((uint64_t)max_seq_id >> 6) + 1);
    for (auto &w : seq_bits) w.store(0, std::memory_order_relaxed);

    /* ---- extract_kmers (src/signature_build.tcc:47-70) ---- */
    const double t0 = now_s();
    std::vector<uint64_t> n_ins(n_threads, 0);
    std::vector<std::vector<uint32_t>> swf(n_threads);
    auto extract_range = [&](int tid, uint64_t lo, uint64_t hi) {
        std::vector<uint32_t> &mine = swf[tid];
        mine.assign(SIGK_N_FUNCTION_SLOTS, 0);
        uint64_t cnt = 0;
        for (uint64_t i = lo; i < hi; ++i) {
            mine[p->function_index[i]]++;                       /* :160 seqs_with_func */
            cnt += load_sequence(t, base[i], p->residues + p->starts[i],
                                 p->starts[i + 1] - p->starts[i], p->function_index[i], p->seq_id[i]);
        }
        n_ins[tid] = cnt;
    };
    if (n_threads < 2) extract_range(0, 0, np);                 /* serial branch :50-56 */
    else {
        std::vector<std::thread> th;
        for (int k = 0; k < n_threads; ++k)
            th.emplace_back(extract_range, k, np * k / n_threads, np * (k + 1) / n_threads);
        for (auto &x : th) x.join();
    }
    for (int k = 0; k < n_threads; ++k) {
        o->n_occurrences += n_ins[k];
        for (int f = 0; f < SIGK_N_FUNCTION_SLOTS; ++f) o->seqs_with_func[f] += swf[k][f];
    }
    const double t1 = now_s();

    /* ---- process_kmers (src/signature_build.tcc:183-213) ---- */
    std::vector<std::vector<Row>> kept(n_threads);
    std::vector<Stats> stats(n_threads);
    const uint64_t n_buckets = t.mask + 1;
    std::atomic<uint64_t> next_chunk(0);
    const uint64_t chunk = 4096;
    auto process_range = [&](int tid) {
        std::vector<const Node *> chain;
        std::vector<unsigned short> offsets;
        std::map<uint16_t, int> func_count;
        for (;;) {
            const uint64_t c0 = next_chunk.fetch_add(chunk);
            if (c0 >= n_buckets) break;
            const uint64_t c1 = std::min(n_buckets, c0 + chunk);
            for (uint64_t b = c0; b < c1; ++b) {
                uint32_t idx = t.heads[b].load(std::memory_order_relaxed);
                if (idx == NIL) continue;
                chain.clear();
                for (; idx != NIL; idx = t.nodes[idx].next) chain.push_back(&t.nodes[idx]);
                /* equal keys contiguous, newest first within a key (walk order kept) */
                if (chain.size() > 1)
                    std::stable_sort(chain.begin(), chain.end(),
                                     [](const Node *a, const Node *b) { return a->kmer < b->kmer; });
                size_t s = 0;
                while (s < chain.size()) {                      /* :194-207 key-change walk */
                    size_t e = s + 1;
                    while (e < chain.size() && chain[e]->kmer == chain[s]->kmer) ++e;
                    stats[tid].n_distinct++;
                    process_kmer_set(chain[s]->kmer, chain.data() + s, e - s, immediate, kept[tid],
                                     stats[tid], seq_bits, offsets, func_count);
                    s = e;
                }
            }
        }
    };
    if (n_threads < 2) process_range(0);
    else {
        std::vector<std::thread> th;
        for (int k = 0; k < n_threads; ++k) th.emplace_back(process_range, k);
        for (auto &x : th) x.join();
    }
    const double t2 = now_s();
    o->t_extract = t1 - t0;
    o->t_process = t2 - t1;

    size_t total = 0;
    for (auto &v : kept) total += v.size();
    o->rows.reserve(total);
    for (int k = 0; k < n_threads; ++k) {
        o->rows.insert(o->rows.end(), kept[k].begin(), kept[k].end());
        std::vector<Row>().swap(kept[k]);
        o->n_distinct += stats[k].n_distinct;
        o->distinct_signatures += stats[k].distinct_signatures;
        for (int f = 0; f < SIGK_N_FUNCTION_SLOTS; ++f) o->distinct_functions[f] += stats[k].distinct_functions[f];
    }
    for (auto &w : seq_bits) o->n_seqs_sig += (uint64_t)__builtin_popcountll(w.load(std::memory_order_relaxed));
    /* Presentation only (the reference iterates in TBB hash order): the order include/sigk.h defines for the table
     * — k-mers without a lower-case residue first, in byte order, then the others by (case-folded bytes, case mask
     * with residue j in bit j). */
    if (!(flags & SIGK_ORACLE_NO_SORT))
        std::sort(o->rows.begin(), o->rows.end(), [](const Row &a, const Row &b) {
            const uint64_t CASE = 0x2020202020202020ull;
            const uint64_t ma = a.kmer & CASE, mb = b.kmer & CASE;
            if ((ma != 0) != (mb != 0)) return mb != 0;
            const uint64_t fa = a.kmer & ~CASE, fb = b.kmer & ~CASE;
            if (fa != fb) return fa < fb;
            return __builtin_bswap64(ma) < __builtin_bswap64(mb);
        });

    const size_t n = o->rows.size();
    o->kmer_bytes.resize(n * 8);
    o->avg.resize(n); o->func.resize(n); o->mean.resize(n); o->median.resize(n); o->var.resize(n);
    for (size_t i = 0; i < n; ++i) {
        const Row &r = o->rows[i];
        for (int j = 0; j < 8; ++j) o->kmer_bytes[i * 8 + j] = (char)((r.kmer >> (56 - 8 * j)) & 0xFF);
        o->avg[i] = r.avg_from_end; o->func[i] = r.function_index;
        o->mean[i] = r.mean; o->median[i] = r.median; o->var[i] = r.var;
    }
    std::vector<Row>().swap(o->rows);
    *out = o;
    return SIGK_OK;
}

int sigk_oracle_result(sigk_oracle *o, sigk_table *out) {
    if (!o || !out) return SIGK_E_INVALID;
    out->n_kept = o->avg.size();
    out->kmer = o->kmer_bytes.data();
    out->avg_from_end = o->avg.data();
    out->function_index = o->func.data();
    out->mean = o->mean.data();
    out->median = o->median.data();
    out->var = o->var.data();
    out->n_occurrences = o->n_occurrences;
    out->n_distinct_kmers = o->n_distinct;
    out->distinct_signatures = o->distinct_signatures;
    out->num_seqs_with_a_signature = o->n_seqs_sig;
    out->distinct_functions = o->distinct_functions.data();
    out->seqs_with_func = o->seqs_with_func.data();
    /* rows of the first section of the table order (k-mers without a lower-case residue) */
    uint64_t upper = 0;
    for (size_t i = 0; i < o->avg.size(); ++i) {
        bool lower = false;
        for (int j = 0; j < 8; ++j) lower = lower || (o->kmer_bytes[i * 8 + j] & 0x20);
        upper += lower ? 0 : 1;
    }
    out->n_upper = upper;
    return SIGK_OK;
}

double sigk_oracle_seconds(const sigk_oracle *o) { return o ? o->t_extract + o->t_process : 0.0; }
double sigk_oracle_extract_seconds(const sigk_oracle *o) { return o ? o->t_extract : 0.0; }
void sigk_oracle_free(sigk_oracle *o) { delete o; }

void sigk_oracle_accumulate(const uint32_t *samples, uint64_t n, int flags, uint16_t *mean,
                            uint16_t *median, uint16_t *var, double *median_f64, double *var_f64) {
    BoostAcc acc;
    acc.immediate = (flags & SIGK_ORACLE_IMMEDIATE_MEAN) != 0;
    for (uint64_t i = 0; i < n; ++i) acc.push(samples[i]);
    if (mean) *mean = n ? u16_from_double(acc.mean_f64()) : 0;
    if (median) *median = u16_from_double(acc.median_f64());
    if (var) *var = u16_from_double(acc.var);
    if (median_f64) *median_f64 = acc.median_f64();
    if (var_f64) *var_f64 = acc.var;
}

uint16_t sigk_oracle_u16_from_double(double d) { return u16_from_double(d); }

int sigk_oracle_keep(int best_count, int count) {
    const float thresh = float(count) * 0.8f;
    return ((float)best_count < thresh) ? 0 : 1;
}

uint64_t sigk_oracle_tbb_hash(const char kmer[8]) {
    size_t h = 0;
    for (int i = 0; i < 8; ++i) h = (h * 17) ^ (unsigned int)kmer[i];
    return (uint64_t)h;
}

}  // extern "C"
