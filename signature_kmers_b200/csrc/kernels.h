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
};

inline uint64_t encode_tiles(uint64_t total_res) { return (total_res + ENC_TILE - 1) / ENC_TILE; }

// scan_state: encode_tiles() u64 words, zeroed; ticket: one zeroed u32; n_out: u64.
cudaError_t launch_encode(const EncodeArgs &a, uint64_t *keys, uint32_t *vals, uint64_t *scan_state,
                          uint32_t *ticket, uint64_t *n_out, cudaStream_t stream);

// seqs_with_func[f]++ per protein (src/signature_build.tcc:160) and len[i] = starts[i+1]-starts[i].
cudaError_t launch_protein_meta(const uint64_t *starts, const uint16_t *func, uint32_t n_prot,
                                uint32_t *len_out, uint32_t *seqs_with_func, cudaStream_t stream);

// ---- stage 2: onesweep LSD radix sort (onesweep.cu) ------------------------
constexpr int SORT_MAX_PASSES = 8;
struct PassPlan {
    int npass;
    int lo[SORT_MAX_PASSES];
    int bits[SORT_MAX_PASSES];
};
PassPlan make_pass_plan(int bit_lo, int bit_hi);

constexpr int OS_THREADS = 512;
constexpr int OS_ITEMS = 15;
constexpr int OS_TILE = OS_THREADS * OS_ITEMS;    // 7680 records per CTA
inline uint64_t onesweep_tiles(uint64_t capacity) { return (capacity + OS_TILE - 1) / OS_TILE; }
// bytes of look-back state one pass needs for `capacity` records
size_t onesweep_lookback_bytes(uint64_t capacity);

// hist: [npass][256] u64, zeroed.  Counts the digit of every pass in one read of the keys.
cudaError_t launch_histogram(const uint64_t *keys, const uint64_t *n_ptr, uint64_t capacity, const PassPlan &plan,
                             uint64_t *hist, int sm_count, cudaStream_t stream);
// bin_base[p][d] = exclusive scan over d of hist[p][d]
cudaError_t launch_scan_bins(const uint64_t *hist, uint64_t *bin_base, int npass, cudaStream_t stream);
// One stable scatter pass on key bits [bit_lo, bit_lo+nbits).  lookback zeroed, ticket zeroed.
cudaError_t launch_onesweep_pass(const uint64_t *keys_in, const uint32_t *vals_in, uint64_t *keys_out,
                                 uint32_t *vals_out, const uint64_t *n_ptr, uint64_t capacity, int bit_lo, int nbits,
                                 const uint64_t *bin_base, void *lookback, uint32_t *ticket, cudaStream_t stream);
cudaError_t onesweep_configure();   // opt in to the dynamic shared memory the pass kernel needs

// ---- stages 3+4: segment reduce, keep/reject, compaction (reduce.cu) -------
struct ProteinMeta {
    const uint16_t *func;      // [n_prot_global]
    const uint32_t *len;       // [n_prot_global]
    const uint32_t *seq_id;    // [n_prot_global]
};

struct KeptColumns {           // device, capacity rows each
    uint64_t *kmer;            // 8 ASCII bytes per row
    uint16_t *avg_from_end, *function_index, *mean, *median, *var;
};

constexpr int SEG_TILE = 4096;     // records per CTA in the head scan
constexpr int CMP_TILE = 2048;     // segments per CTA in the compaction
inline uint64_t seg_tiles(uint64_t capacity) { return (capacity + SEG_TILE - 1) / SEG_TILE; }
inline uint64_t cmp_tiles(uint64_t capacity) { return (capacity + CMP_TILE - 1) / CMP_TILE; }

// Run-length pass: seg_start[s] = index of the first record of k-mer group s.
cudaError_t launch_segment_heads(const uint64_t *keys, const uint64_t *n_ptr, uint64_t capacity, uint32_t *seg_start,
                                 uint64_t *scan_state, uint32_t *ticket, uint64_t *n_seg_out, cudaStream_t stream);
// Per-group tally, 80 % rule, offset median, length statistics (one thread per group).
cudaError_t launch_segment_process(const uint64_t *keys, const uint32_t *vals, const uint64_t *n_ptr,
                                   const uint32_t *seg_start, const uint64_t *n_seg_ptr, uint64_t capacity,
                                   ProteinMeta meta, int order_stats, uint4 *seg_rows, uint32_t *seq_bitmap,
                                   uint32_t *distinct_functions, cudaStream_t stream);
// Keep/compact: kept groups -> table columns in k-mer order.
cudaError_t launch_compact_rows(const uint4 *seg_rows, const uint64_t *n_seg_ptr, uint64_t capacity, KeptColumns out,
                                uint64_t *scan_state, uint32_t *ticket, uint64_t *n_kept_out, cudaStream_t stream);
cudaError_t launch_popcount(const uint32_t *bitmap, uint64_t n_words, uint64_t *out, cudaStream_t stream);

}  // namespace sigk
