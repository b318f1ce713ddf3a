// gen_synth — deterministic synthetic protein sets in the reference's input
// conventions.
//
// Two outputs from the same generator state:
//   * the on-disk tree kmers-build-signatures reads: Seqs/<genome> FASTA
//     (">id\nSEQ\n", scripts/kmers-setup-build.pl:170) and Annotations/0/<genome>
//     ("id\tfunction", :250);
//   * the packed arrays of struct sigk_proteins (include/sigk.h) in canonical
//     order, i.e. exactly what the host side of the drop-in hands the GPU after
//     FunctionMap's gates — used by bench.py and the tests without touching disk.
//
// Every protein is a pure function of (seed, global protein index), so any
// rank can generate any slice and threads need no coordination.
//
// Build: tools/Makefile -> tools/libsigk_synth.so and tools/gen_synth.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>
#include <sys/stat.h>

namespace {

struct Rng {    // splitmix64 stream
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
};
inline uint64_t mix(uint64_t a, uint64_t b) {
    Rng r(a * 0x9E3779B97F4A7C15ULL ^ (b + 0x632BE59BD9B4E019ULL));
    r.next();
    return r.next();
}

const char AA[21] = "ARNDCQEGHILKMFPSTWYV";
// UniProtKB/Swiss-Prot background composition, per mille
const int AAF[20] = {83, 55, 41, 55, 14, 39, 68, 71, 23, 60, 97, 58, 24, 39, 47, 66, 53, 11, 29, 67};

struct Sampler {
    uint8_t table[1024];
    Sampler() {
        int total = 0;
        for (int f : AAF) total += f;
        int k = 0;
        for (int i = 0; i < 20; ++i) {
            int cnt = (int)std::lround(1024.0 * AAF[i] / total);
            for (int j = 0; j < cnt && k < 1024; ++j) table[k++] = (uint8_t)AA[i];
        }
        while (k < 1024) table[k++] = 'L';
    }
    uint8_t draw(Rng &r) const { return table[r.next() >> 54]; }
};
const Sampler g_sampler;

}  // namespace

extern "C" {

struct sigk_synth_params {
    uint64_t n_proteins;
    uint32_t n_functions;
    uint32_t n_genomes;
    uint64_t seed;
    double zipf_s;          // 0 = uniform family sizes
    double mut_rate;        // per-residue substitution rate of a member
    double indel_frac;      // fraction of members carrying one indel
    double x_rate;          // 'X' injection rate
    double lower_rate;      // lower-case injection rate
    double rare_rate;       // B / Z / U / * injection rate
    double domain_frac;     // fraction of families sharing a 40-aa domain
    uint32_t min_reps;      // genomes a function needs to be kept (reference default 3)
    uint32_t max_seqs_per_file;   // 100000 (src/kmers-build-signatures.cc:18)
};

}  // extern "C"

namespace {

struct Synth {
    sigk_synth_params P;
    std::vector<uint64_t> fam_base;          // F+1: first global protein index of each family
    std::vector<uint32_t> anc_start;         // F+1 offsets into anc
    std::vector<uint8_t> anc;                // concatenated ancestors
    std::vector<uint16_t> fam_func;          // function_index per family (0xFFFF = not kept / sentinel)
    std::vector<std::string> index_names;    // function.index order
    uint32_t hypothetical_index = 0xFFFF;
    uint32_t n_kept_functions = 0;
    // canonical order: genome 0's proteins (global idx g, g+G, ...), then genome 1's, ...
    std::vector<uint64_t> genome_first;      // G+1: canonical position (ungated) of each genome's first protein

    uint32_t family_of(uint64_t idx) const {
        return (uint32_t)(std::upper_bound(fam_base.begin(), fam_base.end(), idx) - fam_base.begin() - 1);
    }
    uint64_t global_index(uint64_t canon) const {       // ungated canonical position -> global protein index
        const uint32_t g = (uint32_t)(std::upper_bound(genome_first.begin(), genome_first.end(), canon) - genome_first.begin() - 1);
        return (uint64_t)g + (canon - genome_first[g]) * P.n_genomes;
    }
    static std::string function_name(uint32_t f) { return "Family function " + std::to_string(f); }

    // Generates member `idx`; returns its length.  out may be null (length only).
    uint32_t make_protein(uint64_t idx, uint8_t *out) const {
        const uint32_t f = family_of(idx);
        const uint8_t *a = anc.data() + anc_start[f];
        const uint32_t L = anc_start[f + 1] - anc_start[f];
        Rng r(mix(P.seed ^ 0xABCDEF, idx));
        // indel decision first so the length is known from two draws
        int del_at = -1, del_len = 0, ins_at = -1, ins_len = 0;
        if (r.uni() < P.indel_frac) {
            const uint32_t len = 1 + r.below(10);
            const uint32_t at = r.below(L);
            if (r.next() & 1) { del_at = (int)at; del_len = (int)std::min<uint32_t>(len, L - at > 8 ? L - at - 8 : 0); }
            else { ins_at = (int)at; ins_len = (int)len; }
        } else { r.below(10); r.below(L); r.next(); }
        const uint32_t outL = L - del_len + ins_len;
        if (!out) return outL;
        uint32_t o = 0;
        for (uint32_t i = 0; i < L; ++i) {
            if ((int)i == ins_at) for (int k = 0; k < ins_len; ++k) out[o++] = g_sampler.draw(r);
            if (del_at >= 0 && (int)i >= del_at && (int)i < del_at + del_len) continue;
            uint8_t c = a[i];
            const double u = r.uni();
            if (u < P.mut_rate) c = g_sampler.draw(r);
            out[o++] = c;
        }
        // ambiguity / case injection on the finished member
        for (uint32_t i = 0; i < outL; ++i) {
            const double u = r.uni();
            if (u < P.x_rate) out[i] = 'X';
            else if (u < P.x_rate + P.lower_rate) out[i] = (uint8_t)(out[i] | 0x20);
            else if (u < P.x_rate + P.lower_rate + P.rare_rate && i > 0) out[i] = (uint8_t)"BZU*"[r.below(4)];
        }
        return outL;
    }
};

Synth *make_synth(const sigk_synth_params &P) {
    if (P.n_functions == 0 || P.n_genomes == 0 || P.n_proteins < P.n_functions) return nullptr;
    Synth *S = new Synth;
    S->P = P;
    const uint32_t F = P.n_functions;
    // family sizes
    std::vector<uint64_t> size(F);
    if (P.zipf_s <= 0) {
        for (uint32_t f = 0; f < F; ++f) size[f] = P.n_proteins / F + (f < P.n_proteins % F ? 1 : 0);
    } else {
        std::vector<double> w(F);
        double tot = 0;
        for (uint32_t f = 0; f < F; ++f) { w[f] = std::pow((double)(f + 1), -P.zipf_s); tot += w[f]; }
        const uint64_t spare = P.n_proteins - 3ull * F;     // every family gets at least 3
        uint64_t used = 0;
        for (uint32_t f = 0; f < F; ++f) { size[f] = 3 + (uint64_t)std::floor(spare * w[f] / tot); used += size[f]; }
        for (uint32_t f = 0; used < P.n_proteins; f = (f + 1) % F) { ++size[f]; ++used; }
    }
    S->fam_base.resize(F + 1);
    S->fam_base[0] = 0;
    for (uint32_t f = 0; f < F; ++f) S->fam_base[f + 1] = S->fam_base[f] + size[f];

    // ancestors: clipped log-normal length, mean ~307
    S->anc_start.resize(F + 1);
    S->anc_start[0] = 0;
    const double sigma = 0.5, mu = std::log(307.0) - sigma * sigma / 2;
    std::vector<uint32_t> alen(F);
    for (uint32_t f = 0; f < F; ++f) {
        Rng r(mix(P.seed ^ 0x1234, f));
        const double u1 = std::max(r.uni(), 1e-12), u2 = r.uni();
        const double z = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
        double L = std::exp(mu + sigma * z);
        L = std::min(3000.0, std::max(50.0, L));
        alen[f] = (uint32_t)L;
        S->anc_start[f + 1] = S->anc_start[f] + alen[f];
    }
    S->anc.resize(S->anc_start[F]);
    const uint32_t n_domains = std::max<uint32_t>(1, F / 20);
    for (uint32_t f = 0; f < F; ++f) {
        Rng r(mix(P.seed ^ 0x5678, f));
        uint8_t *a = S->anc.data() + S->anc_start[f];
        for (uint32_t i = 0; i < alen[f]; ++i) a[i] = g_sampler.draw(r);
        if (r.uni() < P.domain_frac && alen[f] >= 80) {      // shared 40-aa domain
            Rng d(mix(P.seed ^ 0x9ABC, r.below(n_domains)));
            const uint32_t at = r.below(alen[f] - 40);
            for (uint32_t i = 0; i < 40; ++i) a[at + i] = g_sampler.draw(d);
        }
    }

    // kept functions and their indices: FunctionMap::process_kept_functions
    // (reference src/function_map.h:257-332): >= min_reps distinct genomes, always
    // "hypothetical protein", ids in std::set<std::string> order, unsigned short counter.
    std::vector<std::pair<std::string, int64_t>> kept;      // name, family (-1 = hypothetical)
    for (uint32_t f = 0; f < F; ++f) {
        const uint64_t genomes = std::min<uint64_t>(size[f], P.n_genomes);
        if (genomes >= P.min_reps) kept.emplace_back(Synth::function_name(f), (int64_t)f);
    }
    kept.emplace_back("hypothetical protein", -1);
    std::sort(kept.begin(), kept.end());
    S->fam_func.assign(F, 0xFFFF);
    S->index_names.clear();
    unsigned short next = 0;
    for (auto &k : kept) {
        const unsigned short id = next++;                   // wraps at 65536 like the reference (:324-330)
        if (k.second >= 0) S->fam_func[(size_t)k.second] = id;
        else S->hypothetical_index = id;
        S->index_names.push_back(k.first);
    }
    S->n_kept_functions = (uint32_t)kept.size();

    const uint32_t G = P.n_genomes;
    S->genome_first.resize(G + 1);
    S->genome_first[0] = 0;
    for (uint32_t g = 0; g < G; ++g) {
        const uint64_t cnt = P.n_proteins > g ? (P.n_proteins - g + G - 1) / G : 0;
        S->genome_first[g + 1] = S->genome_first[g] + cnt;
    }
    return S;
}

}  // namespace

extern "C" {

void sigk_synth_default_params(sigk_synth_params *p) {
    std::memset(p, 0, sizeof *p);
    p->n_proteins = 20000; p->n_functions = 1000; p->n_genomes = 10; p->seed = 1;
    p->zipf_s = 0; p->mut_rate = 0.15; p->indel_frac = 0.05;
    p->x_rate = 1e-3; p->lower_rate = 1e-3; p->rare_rate = 1e-4; p->domain_frac = 0.10;
    p->min_reps = 3; p->max_seqs_per_file = 100000;
}

void *sigk_synth_create(const sigk_synth_params *p) { return p ? make_synth(*p) : nullptr; }
void sigk_synth_destroy(void *h) { delete (Synth *)h; }
uint32_t sigk_synth_kept_functions(void *h) { return ((Synth *)h)->n_kept_functions; }
uint64_t sigk_synth_n_proteins(void *h) { return ((Synth *)h)->P.n_proteins; }

// Lengths of the proteins at ungated canonical positions [lo, hi); gate[i] = 1 if the
// protein passes the reference's gates (its function is kept and its index is not 0xFFFF).
int sigk_synth_lengths(void *h, uint64_t lo, uint64_t hi, uint32_t *len, uint8_t *gate, int n_threads) {
    Synth *S = (Synth *)h;
    if (hi > S->P.n_proteins || lo > hi) return -1;
    if (n_threads < 1) n_threads = 1;
    auto work = [&](int t) {
        const uint64_t a = lo + (hi - lo) * t / n_threads, b = lo + (hi - lo) * (t + 1) / n_threads;
        for (uint64_t c = a; c < b; ++c) {
            const uint64_t idx = S->global_index(c);
            len[c - lo] = S->make_protein(idx, nullptr);
            gate[c - lo] = S->fam_func[S->family_of(idx)] != 0xFFFF;
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
    for (auto &x : th) x.join();
    return 0;
}

// Fills the packed arrays for canonical positions [lo, hi) given starts[] (n_gated+1,
// computed by the caller from sigk_synth_lengths).  seq_id follows the reference:
// file_number * max_seqs_per_file + (proteins with a function string seen so far in the file).
int sigk_synth_fill(void *h, uint64_t lo, uint64_t hi, const uint8_t *gate, const uint64_t *starts, uint8_t *residues,
                    uint16_t *func, uint32_t *seq_id, int n_threads) {
    Synth *S = (Synth *)h;
    if (hi > S->P.n_proteins || lo > hi) return -1;
    if (n_threads < 1) n_threads = 1;
    std::vector<uint64_t> gated_before(hi - lo + 1, 0);
    for (uint64_t c = lo; c < hi; ++c) gated_before[c - lo + 1] = gated_before[c - lo] + (gate[c - lo] ? 1 : 0);
    auto work = [&](int t) {
        const uint64_t a = lo + (hi - lo) * t / n_threads, b = lo + (hi - lo) * (t + 1) / n_threads;
        for (uint64_t c = a; c < b; ++c) {
            if (!gate[c - lo]) continue;
            const uint64_t o = gated_before[c - lo];
            const uint64_t idx = S->global_index(c);
            S->make_protein(idx, residues + starts[o]);
            func[o] = S->fam_func[S->family_of(idx)];
            const uint32_t g = (uint32_t)(idx % S->P.n_genomes);
            seq_id[o] = g * S->P.max_seqs_per_file + (uint32_t)(c - S->genome_first[g]);
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
    for (auto &x : th) x.join();
    return 0;
}

// Writes <dir>/Seqs/1000000.<g> and <dir>/Annotations/0/1000000.<g> and
// <dir>/function.index.expected (idx \t name) for the tests.
int sigk_synth_write_tree(void *h, const char *dir) {
    Synth *S = (Synth *)h;
    const std::string d(dir);
    mkdir(d.c_str(), 0755);
    mkdir((d + "/Seqs").c_str(), 0755);
    mkdir((d + "/Annotations").c_str(), 0755);
    mkdir((d + "/Annotations/0").c_str(), 0755);
    std::vector<uint8_t> buf(4096);
    for (uint32_t g = 0; g < S->P.n_genomes; ++g) {
        const std::string name = "1000000." + std::to_string(g);
        std::ofstream fa(d + "/Seqs/" + name), an(d + "/Annotations/0/" + name);
        if (!fa || !an) return -1;
        uint64_t n = 0;
        for (uint64_t idx = g; idx < S->P.n_proteins; idx += S->P.n_genomes) {
            const uint32_t L = S->make_protein(idx, nullptr);
            if (buf.size() < L) buf.resize(L);
            S->make_protein(idx, buf.data());
            const std::string id = "fig|" + name + ".peg." + std::to_string(++n);
            fa << ">" << id << "\n";
            fa.write((const char *)buf.data(), L);
            fa << "\n";
            an << id << "\t" << Synth::function_name(S->family_of(idx)) << "\n";
        }
    }
    std::ofstream fi(d + "/function.index.expected");
    for (size_t i = 0; i < S->index_names.size(); ++i) fi << (i & 0xFFFF) << "\t" << S->index_names[i] << "\n";
    return 0;
}

}  // extern "C"

#ifdef SIGK_SYNTH_MAIN
int main(int argc, char **argv) {
    sigk_synth_params p;
    sigk_synth_default_params(&p);
    std::string out;
    for (int i = 1; i + 1 < argc; i += 2) {
        const std::string k = argv[i];
        const char *v = argv[i + 1];
        if (k == "--proteins") p.n_proteins = std::strtoull(v, nullptr, 10);
        else if (k == "--functions") p.n_functions = (uint32_t)std::strtoul(v, nullptr, 10);
        else if (k == "--genomes") p.n_genomes = (uint32_t)std::strtoul(v, nullptr, 10);
        else if (k == "--seed") p.seed = std::strtoull(v, nullptr, 10);
        else if (k == "--zipf") p.zipf_s = std::atof(v);
        else if (k == "--mut") p.mut_rate = std::atof(v);
        else if (k == "--out") out = v;
        else { std::fprintf(stderr, "unknown option %s\n", k.c_str()); return 2; }
    }
    if (out.empty()) {
        std::fprintf(stderr, "usage: gen_synth --proteins N --functions F --genomes G --seed S [--zipf s] [--mut m] --out DIR\n");
        return 2;
    }
    void *h = sigk_synth_create(&p);
    if (!h) { std::fprintf(stderr, "bad parameters\n"); return 2; }
    const int rc = sigk_synth_write_tree(h, out.c_str());
    std::printf("wrote %llu proteins, %u functions (%u kept incl. hypothetical), %u genomes to %s\n",
                (unsigned long long)p.n_proteins, p.n_functions, sigk_synth_kept_functions(h), p.n_genomes, out.c_str());
    sigk_synth_destroy(h);
    return rc ? 1 : 0;
}
#endif
