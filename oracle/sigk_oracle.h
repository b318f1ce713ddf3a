/*
 * sigk_oracle.h — CPU oracle for the signature-generation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call it, and only as the checker
 * or the reported CPU baseline.  libsigk never links or calls it.
 *
 * PARITY STATUS.  The reference (olsonanl/signature_kmers) ships no tests, fixtures or golden
 * vectors for this path and cannot be built as a whole in this image (Boost, TBB, cmph and NuDB
 * are absent).  This file restates src/signature_build.tcc:120-293 of the reference.
 *   PINNED: against the reference's own sources (signature_build.{h,tcc}, function_map.h,
 *     seed_utils.h, fasta_parser.cc) compiled unmodified over the stand-in third-party headers of
 *     oracle/refshim/ into oracle/_ref/libref_signature.so — every row, column and counter on
 *     synthetic and hand-made trees (tests/test_reference_shim.py); by the hand-derived known-answer
 *     tests of tests/test_oracle_kat.py; and by an independent
 *     pure-Python restatement (oracle/oracle_py.py).
 *   UNPINNED against a stock reference run: the published algorithms of the two third-party pieces
 *     the path's arithmetic lives in, restated here and in the stand-ins alike:
 *     - Boost.Accumulators (version unpinned by the reference, Makefile:53):
 *       sum / mean / median (= P-square quantile estimator, p = 0.5) / variance
 *       (the iterative form) on an accumulator_set<unsigned short, ...>;
 *     - TBB <= 2020 concurrent_unordered_multimap insertion order (equal keys
 *       iterate newest-first).
 *
 * The types are shared with include/sigk.h so that tests compare like with like.
 */
#ifndef SIGK_ORACLE_H_
#define SIGK_ORACLE_H_

#include "../include/sigk.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sigk_oracle sigk_oracle;

/* n_threads <= 1: the reference's serial branch (src/signature_build.tcc:50-56),
 * canonical insertion order, all five StoredKmerData fields defined.
 * n_threads  > 1: file-parallel insert / bucket-parallel process like the
 * reference's tbb::parallel_for branches; median/var are then order-dependent
 * exactly as they are in the reference (tier A fields stay exact).
 * flags: SIGK_ORACLE_IMMEDIATE_MEAN selects the alternative reading of the
 * Boost dependency resolution (uncertainty 2).        */
#define SIGK_ORACLE_IMMEDIATE_MEAN 0x1
#define SIGK_ORACLE_NO_SORT        0x2   /* leave rows in table order (timing runs) */

int  sigk_oracle_build(const sigk_proteins *p, int n_threads, int flags, sigk_oracle **out);
int  sigk_oracle_result(sigk_oracle *o, sigk_table *out);
/* seconds spent in extract_kmers + process_kmers (the reference's timed region) */
double sigk_oracle_seconds(const sigk_oracle *o);
double sigk_oracle_extract_seconds(const sigk_oracle *o);
void sigk_oracle_free(sigk_oracle *o);

/* Pieces exported so the KATs can pin them one at a time. */

/* accumulator_set<unsigned short, stats<mean, median, variance>> fed with
 * `n` unsigned-int samples in the given order (src/signature_build.tcc:262-279). */
void sigk_oracle_accumulate(const uint32_t *samples, uint64_t n, int flags,
                            uint16_t *mean, uint16_t *median, uint16_t *var,
                            double *median_f64, double *var_f64);

/* (unsigned short)double as g++ -O3 emits it on x86-64 (cvttsd2si, low 16 bits). */
uint16_t sigk_oracle_u16_from_double(double d);

/* keep/reject of src/signature_build.tcc:250-257: 1 = kept. */
int sigk_oracle_keep(int best_count, int count);

/* tbb_hash<8> of src/kmer_data.h:65-74 */
uint64_t sigk_oracle_tbb_hash(const char kmer[8]);

#ifdef __cplusplus
}
#endif
#endif
