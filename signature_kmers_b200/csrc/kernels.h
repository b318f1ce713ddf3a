// kernels.h — host-callable launchers of the libsigk kernels.  Each launcher
// enqueues on `stream` and returns the cudaError_t of the launch; none of them
// synchronises.  Record counts that are only known on the device travel as
// device pointers (n_ptr), and grids are sized for the capacity.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace sigk {

// ---- stage 1: residue scan + k-mer encode (encode.cu) ----------------------
constexpr int ENC_THREADS = 256;
constexpr int ENC_PPT = 16;                       // window positions per thread
constexpr int ENC_TILE = ENC_THREADS * ENC_PPT;   // 4096 positions per CTA
constexpr int ENC_PAD = 64;                       // zero bytes the residue buffer keeps after its last tile

struct EncodeArgs {
    const uint8_t *res;        // device, zero-padded to round_up(total_res, ENC_TILE) + ENC_PAD
    uint64_t total_res;
    const uint64_t *starts;    // device, n_prot + 1
    uint32_t n_prot;
    uint32_t ordinal_base;     // ordinal of local protein 0 in the whole job (multi-GPU)
    const uint32_t *slice_prot; // device, encode_slices()+1: protein holding the first position of each 512-position slice
    uint32_t *prot_windows;     // device, n_prot, zeroed: valid windows per protein (may be null)
};

inline uint64_t encode_tiles(uint64_t total_res) { return (total_res + ENC_TILE - 1) / ENC_TILE; }
inline uint64_t encode_slices(uint64_t total_res) { return encode_tiles(total_res) * (ENC_THREADS / 32); }
cudaError_t launch_slice_index(const uint64_t *starts, uint32_t n_prot, uint64_t total_res, uint32_t *slice_prot, cudaStream_t stream);
// Digit histograms of the main records (hist rows of SORT_BINS entries, zeroed), the side-record count in
// hist[SORT_RADIX] and — when a.prot_windows is set — every protein's window count, all from the residues.
struct PassPlan;
cudaError_t launch_window_count(const EncodeArgs &a, const PassPlan &plan, uint64_t *hist, int sm_count, cudaStream_t stream);

// Encode and route in one pass (multi-GPU; with n_split == 0 there is one owner and this is the plain encode).
// Owner d's records go, in canonical order, to
// dst_keys[d][0...] / dst_vals[d][0...] (this rank's region on owner d: a local send region, or — with
// peer mappings — memory of GPU d written over NVLink); owner_totals[d] receives their number; *overflow is
// set (and the output is incomplete, but the totals are still exact) if a region is too small.  owner_state: encode_slices() * (n_split + 1)
// zeroed u64 words.
struct EncodeSplitArgs {
    const uint64_t *split_codes;   // device, n_split ascending case-folded codes (code35): owner = number of codes <= the record's
    int n_split;
    uint64_t region_stride;        // records each region can hold
    uint64_t *owner_state;
    uint64_t *owner_totals;        // device, n_split + 1
    uint32_t *overflow;            // device, zeroed
    uint64_t *dst_keys[16];
    uint32_t *dst_vals[16];
};
cudaError_t launch_encode_split(const EncodeArgs &a, const EncodeSplitArgs &sp, uint32_t *ticket, cudaStream_t stream);
// The same contract on the tile machinery of the sort (onesweep.cu): one ~900-record run per owner and tile instead of one
// 64-record run per owner and warp slice.  owner_state: encode_route_state_words() zeroed u64 words.
inline uint64_t encode_route_state_words(uint64_t total_res);
cudaError_t launch_encode_route(const EncodeArgs &a, const EncodeSplitArgs &sp, uint32_t *ticket, int sm_count, cudaStream_t stream);

// ---- stage 2: onesweep LSD radix sort (onesweep.cu) ------------------------
constexpr int SORT_MAX_PASSES = 8;
constexpr int SORT_MAX_SEGMENTS = 16; // the first pass reads up to 16 source regions (one per rank) in place
constexpr int SORT_RADIX_BITS = 9;
constexpr int SORT_RADIX = 1 << SORT_RADIX_BITS;
constexpr int SORT_BINS = SORT_RADIX + 8;   // row stride of hist / bin_base / look-back: digits, the side bin, padding
struct PassPlan {
    int npass;
    int lo[SORT_MAX_PASSES];
    int bits[SORT_MAX_PASSES];
};
PassPlan make_pass_plan(int bit_lo, int bit_hi);

#ifndef SIGK_OS_THREADS          // tuning knobs; the defaults are the measured best (profiles/)
#define SIGK_OS_THREADS 512     // sweep on B200, config2 (gpurun_out/sweep_*, profiles/): 512x15 beats 256x13, 384x14, 512x11, 1024x15
#define SIGK_OS_ITEMS 15
#define SIGK_OS_MIN_BLOCKS 2
#endif
constexpr int OS_THREADS = SIGK_OS_THREADS;
constexpr int OS_ITEMS = SIGK_OS_ITEMS;
constexpr int OS_MIN_BLOCKS = SIGK_OS_MIN_BLOCKS;  // CTAs per SM the pass kernel is sized for
constexpr int OS_TILE = OS_THREADS * OS_ITEMS;    // 7680 records per tile
// the fused encode + first pass works on tiles of whole 512-position slices, one per encoding warp; its staging
// carries one pad slot per 16 records, so the tile is a little smaller than a plain pass's
constexpr int ES_WARPS = OS_THREADS / 32 - (OS_THREADS >= 512 ? 2 : 1);
constexpr int ES_TILE = ES_WARPS * 512;           // window positions per tile (7168 = 14 slices with 512 threads)
constexpr int ES_ITEMS = ES_TILE / OS_THREADS;
static_assert(ES_ITEMS * OS_THREADS == ES_TILE && ES_ITEMS <= OS_ITEMS, "fused tile shape");
inline uint64_t onesweep_tiles(uint64_t capacity) { return (capacity + OS_TILE - 1) / OS_TILE; }
inline uint64_t encode_sort_tiles(uint64_t total_res) { return (total_res + ES_TILE - 1) / ES_TILE; }
inline uint64_t encode_route_state_words(uint64_t total_res) { return (encode_sort_tiles(total_res) + 1) * 32; }
// bytes of look-back state one pass needs for `capacity` records (either kind of tile)
size_t onesweep_lookback_bytes(uint64_t capacity);

// The input of a first pass: up to 16 regions read in order (one per source rank; one region on a single GPU).
struct SortSegments {
    int n;
    uint64_t start[SORT_MAX_SEGMENTS + 1];      // first record index of each region in the concatenation; start[n] = total
    uint32_t tile_start[SORT_MAX_SEGMENTS + 1]; // filled by launch_onesweep_first_pass: first tile of each region
    const uint64_t *keys[SORT_MAX_SEGMENTS];
    const uint32_t *vals[SORT_MAX_SEGMENTS];
};

// hist: [npass][SORT_BINS] u64, zeroed.  Counts the digit of every pass in one read of the keys at keys + *off_ptr
// (off_ptr may be null), *n_ptr records.
cudaError_t launch_histogram(const uint64_t *keys, const uint64_t *n_ptr, const uint64_t *off_ptr, uint64_t capacity,
                             const PassPlan &plan, uint64_t *hist, int sm_count, cudaStream_t stream);
// The same over the regions of a first pass, main records only (mask8 == 0); the others are counted into the side bin of row 0.
cudaError_t launch_histogram_main(const SortSegments &seg, const uint64_t *n_ptr, const PassPlan &plan, uint64_t *hist,
                                  int sm_count, cudaStream_t stream);
// bin_base[p][d] = exclusive scan over d of hist[p][d] (row 0 includes the side bin: bin_base[0][SORT_RADIX] = number of
// main records).  counts (may be null) receives {main records, side records, both}.
cudaError_t launch_scan_bins(const uint64_t *hist, uint64_t *bin_base, int npass, uint64_t *counts, cudaStream_t stream);
// One stable scatter pass on key bits [bit_lo, bit_lo+nbits) of the *n_ptr records at offset *off_ptr (null = 0) of the
// in/out buffers.  lookback zeroed (launch_clear_lookback or a memset), ticket zeroed.
cudaError_t launch_onesweep_pass(const uint64_t *keys_in, const uint32_t *vals_in, uint64_t *keys_out,
                                 uint32_t *vals_out, const uint64_t *n_ptr, const uint64_t *off_ptr, uint64_t capacity, int bit_lo, int nbits,
                                 const uint64_t *bin_base, void *lookback, uint32_t *ticket, int sm_count, cudaStream_t stream);
// The first pass of a build over already encoded records: reads the regions of `seg` in place, sorts the main records
// (mask8 == 0) on their lowest digit into keys_out[0 ..) and appends the others, in input order, behind them (the side run).
// n_ptr (may be null) overrides the total when seg has one region whose size is only known on the device.
cudaError_t launch_onesweep_first_pass(const SortSegments &seg, const uint64_t *n_ptr, uint64_t *keys_out, uint32_t *vals_out,
                                       uint64_t capacity, int bit_lo, int nbits, const uint64_t *bin_base, void *lookback,
                                       uint32_t *ticket, int sm_count, cudaStream_t stream);
// Encode fused with that first pass (single GPU): records go from the residues straight to their place after one radix
// pass; they never exist in HBM in canonical order.  hist/bin_base come from launch_window_count + launch_scan_bins.
cudaError_t launch_encode_sort(const EncodeArgs &a, uint64_t *keys_out, uint32_t *vals_out, int bit_lo, int nbits,
                               const uint64_t *bin_base, void *lookback, uint32_t *ticket, int sm_count, cudaStream_t stream);
// zero the look-back rows a pass over *n_ptr records will use
cudaError_t launch_clear_lookback(void *lookback, const uint64_t *n_ptr, uint64_t capacity, cudaStream_t stream);
cudaError_t onesweep_configure();   // opt in to the dynamic shared memory the pass kernels need

// ---- stages 3+4: segment reduce, keep/reject, compaction, order statistics (reduce.cu)
struct KeptColumns {           // device, capacity rows each
    uint64_t *kmer;            // 8 ASCII bytes per row
    uint16_t *avg_from_end, *function_index, *mean, *median, *var;
};
struct OrderWork { uint32_t row, start, count; };   // kept group whose median/var need the ordered walk

constexpr int RED_BATCH = 2048;     // sorted records per run-length tile
inline uint64_t reduce_batches(uint64_t capacity) { return (capacity + RED_BATCH - 1) / RED_BATCH; }
inline uint64_t squeeze_tiles(uint64_t capacity) { return (capacity + 2047) / 2048; }
// Scratch of the reduce stage, carved out of one zeroed u64 buffer of reduce_scan_entries() words
// (reduce_scratch() in reduce.cu): per run-length tile its {heads, listed groups} pair, the scanned bases,
// the group left open at the tile end and the tile's first head; per squeeze tile the tombstones in it
// and before it.
struct ReduceScratch {
    uint64_t *tile_counts, *tile_base;
    uint32_t *tile_open, *tile_first;
    uint32_t *rej_tile, *rej_before;
};
inline uint64_t reduce_scan_entries(uint64_t capacity) { return 3 * (reduce_batches(capacity) + 1) + squeeze_tiles(capacity) + 2; }
size_t reduce_group_entries(uint64_t capacity, int sm_count);  // OrderWork entries: groups of 2..32 records
constexpr uint32_t REDUCE_GIANT = 4096;
size_t reduce_giant_entries(uint64_t capacity);
size_t reduce_long_group_entries(uint64_t capacity);           // OrderWork entries: groups of more than 32 records
size_t reduce_work_entries(uint64_t capacity, int sm_count);   // OrderWork entries: groups whose median/var need the ordered walk
size_t reduce_long_work_entries(uint64_t capacity);            // ... of those, the ones a whole warp walks
cudaError_t reduce_configure(int meta_shift, int rej_shift);   // meta_shift: SIGK_TEST_META_SPREAD (0 in production)

// What a record's protein ordinal is looked up for.  x = protein_length, y = function_index: 8 bytes per
// protein, or — when every protein of the job is shorter than 65 535 residues — 4 bytes, length | function << 16
// (compact), so that the job-wide table stays in L2 as long as possible (2 M proteins = 8 MB).
using ProtMeta = uint2;
struct MetaTable { void *p; bool compact; int shift; };     // shift: entries 2^shift apart (SIGK_TEST_META_SPREAD; 0 in production)
inline size_t meta_bytes(uint64_t n_proteins, bool compact) { return (size_t)n_proteins * (compact ? sizeof(uint32_t) : sizeof(ProtMeta)); }
// meta[first + i] for the n_prot local proteins; seqs_with_func[f]++ (src/signature_build.tcc:160)
cudaError_t launch_protein_meta(const uint64_t *starts, const uint16_t *func, const uint32_t *seq_id, uint32_t n_prot,
                                MetaTable meta, uint64_t first, uint32_t *seqs_with_func, cudaStream_t stream);

// device lists and counters of the segment reduce (counters zeroed before the launch)
struct ReduceLists {
    OrderWork *groups;      uint32_t *n_groups, *next_group;
    OrderWork *long_groups; uint32_t *n_long, *next_long;
    OrderWork *work;        uint32_t *n_work;
    OrderWork *work_long;   uint32_t *n_work_long;
    OrderWork *giant;       uint32_t *n_giant, *next_giant;      // groups of more than REDUCE_GIANT records: a whole CTA each
};
// Run-length + per-group reduce + keep/reject over the sorted records (head_tile_kernel count and emit
// passes, then group_reduce_kernel): one packed row per group in k-mer order (rejected groups
// leave a tombstone), plus the lists of groups whose median/var need the ordered walk.
// scratch_words: reduce_scan_entries() zeroed words, shared with launch_squeeze_rows of the same build.
cudaError_t launch_segment_reduce(const uint64_t *keys, const uint32_t *vals, const uint64_t *n_ptr, uint64_t capacity,
                                  MetaTable meta, uint4 *rows, const ReduceLists &l, uint32_t *prot_rejected,
                                  uint64_t *scratch_words, uint64_t *n_seg_out, int order_stats, int sm_count, cudaStream_t stream,
                                  cudaEvent_t after_count = nullptr, cudaEvent_t after_emit = nullptr);
// distinct_functions[f] += kept rows whose function_index is f (src/signature_build.tcc:286), from the finished
// table's column; max_function = largest function index any protein of the job carries.
cudaError_t launch_function_histogram(const uint16_t *function_index, const uint64_t *n_kept_ptr, uint64_t capacity,
                                      uint32_t max_function, uint32_t *distinct_functions, int sm_count, cudaStream_t stream);
// median / var of the groups listed in `work`, patched into their rows.
// ---- fasta.cu: FASTA bytes -> record table + packed residues ----
constexpr uint32_t FASTA_TILE = 8192;           // bytes per tile; a tile never spans two files
struct FastaTile { uint64_t begin; uint32_t n; uint32_t file_start; };
struct FastaOut {
    uint8_t *residues = nullptr;                // sequence characters of all records, in file order
    uint64_t *header_pos = nullptr, *id_end = nullptr, *line_end = nullptr, *seq_begin = nullptr;    // per record
    uint64_t *err_pos = nullptr;                // position | state << 60 of the reported characters
    uint32_t *err_record = nullptr;
    uint64_t err_capacity = 0;
};
constexpr uint32_t FASTA_CHUNKS_PER_TILE = FASTA_TILE / 16;      // a thread's 16 bytes
cudaError_t launch_fasta_tile_functions(const uint8_t *bytes, const FastaTile *tiles, uint32_t n_tiles, uint32_t *chunk_before, uint32_t *tile_fn,
                                        uint8_t *tile_state, cudaStream_t stream);
cudaError_t launch_fasta_count(const uint8_t *bytes, const FastaTile *tiles, uint32_t n_tiles, const uint8_t *tile_state, const uint32_t *chunk_before,
                               uint16_t *chunk_counts, uint64_t *tile_packed, uint64_t *tile_prefix, uint64_t *totals, cudaStream_t stream);
cudaError_t launch_fasta_emit(const uint8_t *bytes, const FastaTile *tiles, uint32_t n_tiles, const uint8_t *tile_state, const uint32_t *chunk_before,
                              uint16_t *chunk_counts, uint64_t *tile_prefix, const FastaOut &out, cudaStream_t stream);
cudaError_t launch_fasta_gather(const uint8_t *stream_bytes, const uint64_t *src_begin, const uint64_t *starts, uint32_t n_proteins,
                                uint8_t *residues, cudaStream_t stream);

cudaError_t launch_ddiv_check(const double *a, const double *b, uint64_t n, double *inl, double *lib, cudaStream_t stream);
cudaError_t launch_order_stats(const uint32_t *vals, MetaTable meta, const OrderWork *work, const uint32_t *n_work,
                               uint32_t *next_work, const OrderWork *work_long, const uint32_t *n_work_long, uint32_t *next_long,
                               uint64_t capacity, uint4 *rows, int sm_count, cudaStream_t stream);
// Compaction: kept rows -> table columns (tombstones dropped, order kept).  scratch_words: the buffer the
// segment reduce of this build used (it holds the per-tile tombstone counts).
cudaError_t launch_squeeze_rows(const uint4 *rows, const uint64_t *n_seg_ptr, uint64_t capacity, KeptColumns out,
                                uint64_t *scratch_words, uint64_t *n_kept_out, uint64_t *n_side_kept /* zeroed */, cudaStream_t stream);
// bitmap bit seq_id[i] is set iff protein i has more occurrences (prot_windows, from encode) than occurrences in
// rejected groups (prot_rejected, from the reduce): kmer_stats_.seqs_with_a_signature, src/signature_build.tcc:274
cudaError_t launch_signature_flags(const uint32_t *prot_windows, const uint32_t *prot_rejected, const uint32_t *seq_id,
                                   uint32_t n_prot, uint32_t *bitmap, cudaStream_t stream);
cudaError_t launch_popcount(const uint32_t *bitmap, uint64_t n_words, uint64_t *out, cudaStream_t stream);

// ---- consumer side: batch lookups against the resident table (lookup.cu) ----
// rows[g] = table row of the window starting at residue position g (for_each_kmer + KeptKmerDB::fetch), else 0xFFFFFFFF.
// table_kmers: the table's k-mer column (8 ASCII bytes per row, sorted by bytes), *n_rows_ptr rows.
cudaError_t launch_lookup(const uint8_t *res, const uint64_t *starts, uint32_t n_prot, uint64_t total, const uint64_t *table_kmers,
                          const uint64_t *n_rows_ptr, uint32_t *rows, cudaStream_t stream);

}  // namespace sigk
