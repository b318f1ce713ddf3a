/*
 * sigk.h — C ABI of libsigk, the B200 (sm_100a) signature k-mer builder.
 *
 * This is the drop-in boundary for the signature-generation hot path of
 * olsonanl/signature_kmers.  One `sigk_handle` replaces the two calls the
 * reference's main() makes on its SignatureBuilder<8>
 *
 *      builder.extract_kmers(deleted_fids);     src/kmers-build-signatures.cc:194
 *      builder.process_kmers();                 src/kmers-build-signatures.cc:196
 *
 * (bodies: src/signature_build.tcc:47-293) and the three accessors main() then
 * reads: kept_kmers() / kmer_stats() (src/signature_build.h:106-107).
 *
 * The host keeps everything the reference does on strings (FASTA parsing,
 * FunctionMap, the per-protein gates of signature_build.tcc:120-160) and hands
 * the library packed proteins in canonical order; the library does everything
 * from window enumeration onward on the GPU and returns the kept-k-mer table
 * with the StoredKmerData fields of src/kmer_data.h:114-128.
 *
 * Plain C types only.  Every call returns 0 on success or a negative SIGK_E_*
 * code; sigk_last_error() gives the text.  One handle = one host thread = one
 * GPU.  There is no CPU fallback: without a usable CUDA device every compute
 * entry point fails with SIGK_E_CUDA.
 */
#ifndef SIGK_H_
#define SIGK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIGK_ABI_VERSION 2

#define SIGK_K 8                    /* src/kmers-build-signatures.cc:17 (const int K = 8) */
#define SIGK_UNDEFINED_FUNCTION 0xFFFFu   /* src/kmer_data.h:23 */
#define SIGK_N_FUNCTION_SLOTS 65536

enum {
    SIGK_OK = 0,
    SIGK_E_INVALID = -1,    /* bad argument / call order */
    SIGK_E_CUDA = -2,       /* CUDA runtime or driver error, or no device */
    SIGK_E_NOMEM = -3,      /* host or device allocation failed */
    SIGK_E_UNSUPPORTED = -4,
    SIGK_E_COMM = -5        /* multi-GPU exchange failed */
};

typedef struct sigk_handle sigk_handle;

/* Replaces SignatureBuilder<K>::SignatureBuilder(n_threads, max_seqs_per_file)
 * (src/signature_build.tcc:3-8): n_threads has no meaning on the GPU; the
 * seq_id arithmetic that used max_seqs_per_file stays on the host.          */
typedef struct sigk_config {
    int32_t abi_version;    /* SIGK_ABI_VERSION */
    int32_t k;              /* must be 8 */
    int32_t device;         /* CUDA device ordinal */
    int32_t rank;           /* this process's rank in the k-mer range partition (0 if single GPU) */
    int32_t world;          /* number of ranks (1 if single GPU) */
    uint32_t flags;         /* SIGK_F_* */
} sigk_config;

#define SIGK_F_NO_ORDER_STATS 0x1u  /* skip the order-dependent median/var columns (tier B); they read 0 */

/* Packed input, canonical order (= the order load_kmers_from_sequence would
 * have been called with --n-threads 1: files in all_fasta_data_ order,
 * proteins in file order; src/signature_build.tcc:50-56).  Only proteins that
 * passed the reference's gates (non-empty id, not deleted, function string
 * present, function kept: src/signature_build.tcc:94,122-158) are listed.
 *
 *   residues        concatenated raw sequence bytes exactly as FastaParser
 *                   delivered them (case preserved), protein i occupying
 *                   [starts[i], starts[i+1]); no separators needed
 *   starts          n_proteins+1 offsets, starts[0] == 0, non-decreasing
 *   function_index  per protein, != 0xFFFF (FunctionMap::lookup_index)
 *   seq_id          per protein, the reference's seq_id
 *                   (file_number*max_seqs_per_file + k; :91,:138)
 *
 * The arrays must stay valid and unchanged until sigk_build/sigk_upload
 * returns; use sigk_host_alloc for them to get pinned (DMA-able) memory.   */
typedef struct sigk_proteins {
    const uint8_t  *residues;
    const uint64_t *starts;
    const uint16_t *function_index;
    const uint32_t *seq_id;
    uint64_t n_proteins;
} sigk_proteins;

/* Kept-k-mer table: one row per KeptKmer<8> (src/signature_build.h:34-42),
 * struct-of-arrays, in two sorted sections (the reference's own order is TBB
 * hash order and no consumer depends on it):
 *   rows [0, n_upper)        k-mers without a lower-case residue, ascending
 *                            unsigned-char order of their 8 bytes;
 *   rows [n_upper, n_kept)   k-mers with at least one lower-case residue (the
 *                            reference keeps case: 'a' != 'A',
 *                            src/signature_build.tcc:167), ascending by
 *                            (upper-cased bytes, then the case mask: bit j set
 *                            iff residue j is lower case).
 * A lookup is a binary search in the section the query's case pattern picks
 * (sigk_lookup, SortedKmerDb).
 * Columns are the StoredKmerData fields (src/kmer_data.h:114-128).  Pointers
 * are host memory owned by the handle, valid until the next build or destroy.
 * The three counters are the ones process_kmers prints
 * (src/signature_build.tcc:210-212); distinct_functions / seqs_with_func are
 * KmerStatistics (src/signature_build.h:44-50) as dense [65536] arrays.      */
typedef struct sigk_table {
    uint64_t n_kept;                    /* "Kept N kmers" == distinct_signatures */
    const char     *kmer;               /* n_kept * 8 raw ASCII bytes */
    const uint16_t *avg_from_end;
    const uint16_t *function_index;
    const uint16_t *mean;
    const uint16_t *median;
    const uint16_t *var;
    uint64_t n_occurrences;             /* valid windows inserted (multimap size) */
    uint64_t n_distinct_kmers;          /* groups seen by process_kmers */
    uint64_t distinct_signatures;
    uint64_t num_seqs_with_a_signature;
    const uint32_t *distinct_functions; /* [65536] kept k-mers per function */
    const uint32_t *seqs_with_func;     /* [65536] proteins per function (:160) */
    uint64_t n_upper;                   /* rows of the first section (see above); ignored by sigk_set_table */
} sigk_table;

/* Device time of the last build, CUDA events on the library's own stream. */
typedef struct sigk_timings {
    float h2d_ms;
    float encode_ms;        /* everything before the first radix pass: protein table, window count (one GPU) or encode + route (multi-GPU) */
    float histogram_ms;     /* digit histogram of already encoded records (multi-GPU and the unfused path) */
    float sort_ms;          /* all onesweep passes, the fused encode + first pass included, and the side run's sort */
    float reduce_ms;        /* giant pre-pass + streaming segment reduce + keep/reject */
    float order_stats_ms;   /* median/var worklist */
    float squeeze_ms;       /* ordered compaction of kept rows into the table columns */
    float exchange_ms;      /* multi-GPU partition + all-to-all */
    float d2h_ms;
    float device_total_ms;  /* first kernel start -> last kernel end */
    uint32_t sort_passes;
    uint32_t record_bytes;
    uint32_t key_bytes;
    uint32_t kernel_launches;
    float pass_ms[8];       /* per onesweep pass of the main run; [0] is the first pass (fused with the encode on one GPU) */
    float count_ms;         /* window count pass (digit histograms from the residues) */
    float side_sort_ms;     /* histogram + passes of the side run (records with a lower-case residue) */
    float reduce_comm_ms;   /* multi-GPU: the all-reduce of the per-protein rejected-occurrence counts (inside reduce_ms) */
    float reduce_count_ms;  /* inside reduce_ms: run-length count pass + scan */
    float reduce_emit_ms;   /* inside reduce_ms: run-length emit pass (single-record groups finished) + open groups */
    float reduce_groups_ms; /* inside reduce_ms: groups of two or more records */
    uint64_t records_sorted;      /* records this GPU sorted and reduced (its k-mer range's share with a communicator) */
    uint64_t exchange_bytes_out;  /* multi-GPU: bytes of records this GPU stored into other GPUs' landing zones */
} sigk_timings;

const char *sigk_version(void);
int sigk_device_count(void);

int  sigk_create(const sigk_config *cfg, sigk_handle **out);
void sigk_destroy(sigk_handle *h);   /* collective after sigk_comm_join: every rank calls it (peer mappings are torn down together) */
const char *sigk_last_error(const sigk_handle *h);   /* h may be NULL: last create error */

/* Pinned host memory for input arrays (optional but needed for full PCIe rate). */
void *sigk_host_alloc(size_t bytes);
void  sigk_host_free(void *p);

/* Declare the input; no copy is made. */
int sigk_set_proteins(sigk_handle *h, const sigk_proteins *p);

/* extract_kmers + process_kmers in one call: H2D, all kernels, D2H of the
 * kept table.  Equivalent to upload + build_device + download.             */
int sigk_build(sigk_handle *h);

/* The same, split so that a caller can keep inputs resident in HBM. */
int sigk_upload(sigk_handle *h);
int sigk_build_device(sigk_handle *h);
int sigk_download(sigk_handle *h);

int sigk_result(sigk_handle *h, sigk_table *out);
int sigk_get_timings(const sigk_handle *h, sigk_timings *out);

/* The library enqueues on its own stream; these let a caller bracket any
 * number of sigk_build_device calls with CUDA events on that stream.        */
int sigk_synchronize(sigk_handle *h);
int sigk_event_record(sigk_handle *h, int slot /* 0..3 */);
int sigk_event_elapsed_ms(sigk_handle *h, int slot_a, int slot_b, float *ms);

/* Multi-GPU (one process per GPU): rank 0 makes an id, the launcher ships the
 * 128 bytes to every rank, every rank joins.  After that sigk_build routes
 * each record to the rank that owns its k-mer range (one all-to-all) and
 * returns this rank's slice of the kept table (k-mer ranges are cut on the
 * case-folded k-mer, ascending in rank): first sections concatenated in rank
 * order, then second sections in rank order, are the whole table in order.  */
#define SIGK_COMM_ID_BYTES 128
int sigk_comm_make_id(void *id128);
int sigk_comm_join(sigk_handle *h, const void *id128);

/* ---- stand-alone kernels, exported for the parity tests -------------------
 * Each runs one stage on host arrays through a temporary device copy.       */

/* Stage 1: window validity + 43-bit group code (base-20 code of the case-folded residues << 8 | case mask).
 * out_code[n_windows], out_ordinal[n_windows], out_offset[n_windows];
 * returns the number of valid windows through *n_out.                       */
int sigk_dbg_encode(sigk_handle *h, const sigk_proteins *p,
                    uint64_t *out_code, uint32_t *out_ordinal, uint16_t *out_offset,
                    uint64_t capacity, uint64_t *n_out);

/* Stage 2: stable LSD radix sort of (key,value) pairs on key bits [bit_lo,bit_hi). */
int sigk_dbg_sort_pairs(sigk_handle *h, uint64_t *keys, uint32_t *vals, uint64_t n,
                        int bit_lo, int bit_hi);

/* ---- FASTA bytes -> proteins on the device ------------------------------------
 * Replaces FastaParser::parse_char (src/fasta_parser.h:38-144) as driven by
 * SignatureBuilder<K>::load_kmers_from_fasta (src/signature_build.tcc:84-102): the same five-state machine with
 * every quirk kept ('\r' dropped; anything but '>' before the first header reported; a blank ends the id and
 * starts the definition; letters and '*' are sequence, but '*' at the start of a line is reported and dropped;
 * the '>' that follows a header-only record is reported and that header's letters join the open record).
 *
 * sigk_fasta_parse: `bytes` holds n_files files, file f at bytes[file_begin[f] .. file_begin[f] + file_len[f])
 * with file_begin[f] a multiple of 16, ascending and non-overlapping; every file starts a fresh parser.
 * The records come back in file order: record r has its '>' at header_pos[r]; its id is the bytes after it up to
 * id_end[r], its definition the bytes from id_end[r] up to line_end[r] (both minus any '\r'; a position equal to
 * SIGK_FASTA_NO_POS means "the file ended first"); its sequence is seq_begin[r] .. seq_begin[r+1] of a residue
 * stream that stays on the device.  errors[] lists what the reference would have reported on stderr
 * (position | state << 60, state 0 "Missing >", 3 "Bad data character", 4 "Bad id or data character"; at most
 * SIGK_FASTA_MAX_ERRORS of n_errors).  The arrays belong to the handle and live until the next parse.
 * The reference calls its callback once more per file at the end of input with whatever is pending — the last
 * record, already in this table — and then again with an empty record; callers that mirror the callback add those.
 *
 * sigk_fasta_commit: record r with keep[r] != 0 becomes a protein with function_index[r] and seq_id[r]
 * (what load_kmers_from_sequence decides per id, src/signature_build.tcc:118-160, stays with the caller: it is
 * a string-keyed map lookup).  The kept records' residues are gathered on the device, in record order, into the
 * input of the next sigk_upload / sigk_build, exactly as if sigk_set_proteins had been given them. */
#define SIGK_FASTA_NO_POS 0xFFFFFFFFFFFFFFFFull
#define SIGK_FASTA_MAX_ERRORS 65536
typedef struct sigk_fasta_records {
    uint64_t n_records, n_residues, n_errors;
    const uint64_t *header_pos, *id_end, *line_end;     /* [n_records] */
    const uint64_t *seq_begin;                          /* [n_records + 1] */
    const uint64_t *errors;                             /* [min(n_errors, SIGK_FASTA_MAX_ERRORS)] */
    const uint32_t *error_record;                       /* record open at the error, 0xFFFFFFFF before the first */
    float h2d_ms, parse_ms, d2h_ms;
} sigk_fasta_records;
int sigk_fasta_parse(sigk_handle *h, const uint8_t *bytes, const uint64_t *file_begin, const uint64_t *file_len,
                     uint64_t n_files, sigk_fasta_records *out);
int sigk_fasta_commit(sigk_handle *h, const uint8_t *keep, const uint16_t *function_index, const uint32_t *seq_id);
/* tests: the residue stream of the last parse, n_residues bytes */
int sigk_dbg_fasta_stream(sigk_handle *h, uint8_t *out);

/* The order statistics divide with an inlined IEEE division (csrc/length_acc.cuh); this runs it beside the
 * toolkit's __ddiv_rn on host arrays: inl[i], lib[i] = a[i] / b[i] either way.  They must agree bit for bit. */
int sigk_dbg_ddiv(sigk_handle *h, const double *a, const double *b, uint64_t n, double *inl, double *lib);

/* ---- consumer side of the table --------------------------------------------
 * Batch form of what FunctionCaller::process_aa_seq does per window (src/call_functions.tcc:276-282):
 * for_each_kmer (src/kmer_data.h:76-102: a window is skipped when '*' or 'X' lies in it or right behind it)
 * + KeptKmerDB::fetch (src/kept_kmer_db.h:20-28: exact membership) for every window of every query protein.
 * rows[g] (host, starts[n_proteins] entries) = row of the window starting at residue position g in the table
 * that is resident on the device — the last build's, or the one given to sigk_set_table — or 0xFFFFFFFF.
 * The caller reads avg_from_end / function_index / mean of a hit from its host copy of the table (sigk_result
 * or the file it loaded) and feeds the hits, in position order, to the call logic.  With a communicator each
 * rank looks up in its own slice of the table. */
int sigk_lookup(sigk_handle *h, const uint8_t *residues, const uint64_t *starts, uint64_t n_proteins, uint32_t *rows);
/* Make a table the lookup target without building it here: only t->kmer (8 bytes per row, in the two-section
 * order sigk_result returns them in) and t->n_kept are read. */
int sigk_set_table(sigk_handle *h, const sigk_table *t);

/* 43-bit group code <-> 8 ASCII bytes (host side, no GPU).  Table order = ascending ((code & 0xFF) != 0, code). */
uint64_t sigk_kmer_encode(const char kmer[8]);          /* UINT64_MAX if any residue invalid */
void     sigk_kmer_decode(uint64_t code, char kmer[8]);

#ifdef __cplusplus
}
#endif
#endif /* SIGK_H_ */
