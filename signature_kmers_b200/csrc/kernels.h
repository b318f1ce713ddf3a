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
inline uint64_t encode_scan_entries(uint64_t total_res) { return encode_tiles(total_res) * (ENC_THREADS / 32) + 1; }

// scan_state: encode_scan_entries() u64 words, zeroed; ticket: one zeroed u32; n_out: u64.
cudaError_t launch_encode(const EncodeArgs &a, uint64_t *keys, uint32_t *vals, uint64_t *scan_state,
                          uint32_t *ticket, uint64_t *n_out, cudaStream_t stream);


// Multi-GPU: encode and route in one pass.  Owner d's records go, in canonical order, to
// dst_keys[d][0...] / dst_vals[d][0...] (this rank's region on owner d: a local send region, or — with
// peer mappings — memory of GPU d written over NVLink); owner_totals[d] receives their number; *overflow is
// set (and the output is incomplete) if a region is too small.  owner_state: encode_slices() * (n_split + 1)
// zeroed u64 words.
struct EncodeSplitArgs {
    const uint64_t *split_codes;   // device, n_split ascending codes: owner = number of codes <= the record's code
    int n_split;
    uint64_t region_stride;        // records each region can hold
    uint64_t *owner_state;
    uint64_t *owner_totals;        // device, n_split + 1
    uint32_t *overflow;            // device, zeroed
    uint64_t *dst_keys[16];
    uint32_t *dst_vals[16];
};
cudaError_t launch_encode_split(const EncodeArgs &a, const EncodeSplitArgs &sp, uint32_t *ticket, cudaStream_t stream);

// ---- stage 2: onesweep LSD radix sort (onesweep.cu) ------------------------
constexpr int SORT_MAX_PASSES = 8;
constexpr int SORT_MAX_SPLIT = 15;    // the partition pass routes to at most 16 ranks
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
constexpr int OS_TILE = OS_THREADS * OS_ITEMS;    // 3328 records per CTA
inline uint64_t onesweep_tiles(uint64_t capacity) { return (capacity + OS_TILE - 1) / OS_TILE; }
// bytes of look-back state one pass needs for `capacity` records
size_t onesweep_lookback_bytes(uint64_t capacity);

// hist: [npass][SIGK_RADIX] u64, zeroed.  Counts the digit of every pass in one read of the keys.
cudaError_t launch_histogram(const uint64_t *keys, const uint64_t *n_ptr, uint64_t capacity, const PassPlan &plan,
                             uint64_t *hist, int sm_count, cudaStream_t stream);
// bin_base[p][d] = exclusive scan over d of hist[p][d]
cudaError_t launch_scan_bins(const uint64_t *hist, uint64_t *bin_base, int npass, cudaStream_t stream);
// One stable scatter pass on key bits [bit_lo, bit_lo+nbits).  lookback zeroed, ticket zeroed.
cudaError_t launch_onesweep_pass(const uint64_t *keys_in, const uint32_t *vals_in, uint64_t *keys_out,
                                 uint32_t *vals_out, const uint64_t *n_ptr, uint64_t capacity, int bit_lo, int nbits,
                                 const uint64_t *bin_base, void *lookback, uint32_t *ticket, cudaStream_t stream);
cudaError_t onesweep_configure();   // opt in to the dynamic shared memory the pass kernel needs
// The same kernel as a stable multi-way split: digit = number of split_codes <= the record's k-mer code
// (the rank owning the record).  bin_base[d] = first output slot of rank d (SIGK_RADIX entries).
cudaError_t launch_onesweep_partition(const uint64_t *keys_in, const uint32_t *vals_in, uint64_t *keys_out,
                                      uint32_t *vals_out, const uint64_t *n_ptr, uint64_t capacity,
                                      const uint64_t *split_codes, int n_split, const uint64_t *bin_base, void *lookback,
                                      uint32_t *ticket, cudaStream_t stream);

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
size_t reduce_long_group_entries(uint64_t capacity);           // OrderWork entries: groups of more than 32 records
size_t reduce_work_entries(uint64_t capacity, int sm_count);   // OrderWork entries: groups whose median/var need the ordered walk
size_t reduce_long_work_entries(uint64_t capacity);            // ... of those, the ones a whole warp walks
cudaError_t reduce_configure();

// What a record's protein ordinal is looked up for.  x = protein_length, y = function_index: 8 bytes per
// protein, or — when every protein of the job is shorter than 65 535 residues — 4 bytes, length | function << 16
// (compact), so that the job-wide table stays in L2 as long as possible (2 M proteins = 8 MB).
using ProtMeta = uint2;
struct MetaTable { void *p; bool compact; };
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
};
// Run-length + per-group reduce + keep/reject over the sorted records (head_tile_kernel count and emit
// passes, then group_reduce_kernel): one packed row per group in k-mer order (rejected groups
// leave a tombstone), plus the lists of groups whose median/var need the ordered walk.
// scratch_words: reduce_scan_entries() zeroed words, shared with launch_squeeze_rows of the same build.
cudaError_t launch_segment_reduce(const uint64_t *keys, const uint32_t *vals, const uint64_t *n_ptr, uint64_t capacity,
                                  MetaTable meta, uint4 *rows, const ReduceLists &l, uint32_t *prot_rejected,
                                  uint64_t *scratch_words, uint64_t *n_seg_out, int order_stats, int sm_count, cudaStream_t stream);
// distinct_functions[f] += kept rows whose function_index is f (src/signature_build.tcc:286), from the finished
// table's column; max_function = largest function index any protein of the job carries.
cudaError_t launch_function_histogram(const uint16_t *function_index, const uint64_t *n_kept_ptr, uint64_t capacity,
                                      uint32_t max_function, uint32_t *distinct_functions, int sm_count, cudaStream_t stream);
// median / var of the groups listed in `work`, patched into their rows.
cudaError_t launch_order_stats(const uint32_t *vals, MetaTable meta, const OrderWork *work, const uint32_t *n_work,
                               uint32_t *next_work, const OrderWork *work_long, const uint32_t *n_work_long, uint32_t *next_long,
                               uint64_t capacity, uint4 *rows, int sm_count, cudaStream_t stream);
// Compaction: kept rows -> table columns (tombstones dropped, order kept).  scratch_words: the buffer the
// segment reduce of this build used (it holds the per-tile tombstone counts).
cudaError_t launch_squeeze_rows(const uint4 *rows, const uint64_t *n_seg_ptr, uint64_t capacity, KeptColumns out,
                                uint64_t *scratch_words, uint64_t *n_kept_out, cudaStream_t stream);
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
