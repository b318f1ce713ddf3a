// reduce.cu — stages 3 and 4: run-length the sorted records into k-mer groups,
// reduce each group (function tally, offset median, length statistics), apply
// the reference's keep/reject rule and compact the kept rows into the table.
//
// Replaces SignatureBuilder<K>::process_kmers / process_kmer_set
// (reference src/signature_build.tcc:183-293):
//   tally + arg-max       :203, :228-248   -> majority vote (see below)
//   80 % rule             :250-257         -> same float32 ops
//   offsets median        :273, :281-282   -> rank n/2 of the group's offsets
//   mean / median / var   :262-279         -> Boost.Accumulators restated:
//        sum in unsigned short (wraps), P-square median, iterative variance,
//        items visited newest-first (TBB multimap order inside a key), i.e.
//        backwards through the stably sorted group
//   kept row              :288             -> StoredKmerData columns
//   statistics            :274, :285-286   -> seq bitmap, per-function counts
//
// Why a majority vote is enough for the tally: a group is kept only if
// best_count >= 0.8 * count, so a kept group's best function is a strict
// majority and Boyer-Moore finds it in one pass with O(1) state; if the vote's
// candidate is not a strict majority no function is, the true best_count is
// <= count/2 and the float test rejects for every count.  Ties therefore never
// reach a kept row and the "lowest index wins" order of the reference's std::map
// walk cannot matter.
//
// Floating point: every double operation below is an explicit round-to-nearest
// intrinsic (no FMA contraction), matching g++ -O3 without -march
// (reference Makefile:42-48).  Marker positions are small integers, so they are
// kept as ints and converted where Boost has them as doubles (exact).
#include "kernels.h"
#include "sigk_common.cuh"
#include "length_acc.cuh"

namespace sigk {

namespace {

constexpr int SEG_THREADS = 512;
constexpr int SEG_ITEMS = SEG_TILE / SEG_THREADS;     // 8
constexpr int SEG_WARPS = SEG_THREADS / 32;

// Shared pattern of the two ordered compactions below: every thread holds
// ITEMS flags in warp-striped order (item i of lane l = element wbase + 32 i + l);
// returns the tile-local rank base of this warp and the tile total.
template <int THREADS, int ITEMS>
SIGK_D uint32_t striped_flag_scan(const unsigned (&ball)[ITEMS], uint32_t *s_scan, uint32_t *total) {
    uint32_t warp_total = 0;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) warp_total += __popc(ball[i]);
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t excl = block_exclusive_scan<THREADS>(lane == 0 ? warp_total : 0u, s_scan, total);
    return __shfl_sync(0xffffffffu, excl, 0);
}

__global__ void __launch_bounds__(SEG_THREADS)
segment_heads_kernel(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ n_ptr, uint32_t *__restrict__ seg_start,
                     uint64_t *__restrict__ scan_state, uint32_t *__restrict__ ticket, uint64_t *__restrict__ n_seg_out) {
    __shared__ uint32_t s_scan[SEG_WARPS + 2];
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_base;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint64_t n = *n_ptr;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t tile_start = (uint64_t)tile * SEG_TILE;
    if (tile_start >= n) return;

    unsigned ball[SEG_ITEMS];
    const uint64_t wbase = tile_start + (uint64_t)warp * (SEG_ITEMS * 32);
#pragma unroll
    for (int i = 0; i < SEG_ITEMS; ++i) {
        const uint64_t p = wbase + i * 32 + lane;
        bool head = false;
        if (p < n) {
            const uint64_t c = sigk_key_code(__ldg(keys + p));
            head = (p == 0) || (sigk_key_code(__ldg(keys + p - 1)) != c);      // kmer != cur, tcc:194
        }
        ball[i] = __ballot_sync(0xffffffffu, head);
    }
    uint32_t total;
    uint32_t run = striped_flag_scan<SEG_THREADS, SEG_ITEMS>(ball, s_scan, &total);
    if (tid == 0) {
        const uint64_t base = chained_scan_exclusive(scan_state, tile, total);
        s_base = base;
        if (tile_start + SEG_TILE >= n) *n_seg_out = base + total;
    }
    __syncthreads();
    const uint64_t base = s_base;
#pragma unroll
    for (int i = 0; i < SEG_ITEMS; ++i) {
        if ((ball[i] >> lane) & 1u)
            seg_start[base + run + __popc(ball[i] & ((1u << lane) - 1u))] = (uint32_t)(wbase + i * 32 + lane);
        run += __popc(ball[i]);
    }
}

// seg_rows[s] packing (one uint4 per k-mer group):
//   x = code[31:0]   y = code[42:32] | avg_from_end << 11 | keep << 31
//   z = function_index | mean << 16   w = median | var << 16
__global__ void __launch_bounds__(256)
segment_process_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals,
                       const uint64_t *__restrict__ n_ptr, const uint32_t *__restrict__ seg_start,
                       const uint64_t *__restrict__ n_seg_ptr, ProteinMeta meta, int order_stats,
                       uint4 *__restrict__ seg_rows, uint32_t *__restrict__ seq_bitmap,
                       uint32_t *__restrict__ distinct_functions) {
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t n_seg = *n_seg_ptr;
    if (s >= n_seg) return;
    const uint64_t n = *n_ptr;
    const uint64_t start = seg_start[s];
    const uint64_t end = (s + 1 < n_seg) ? (uint64_t)seg_start[s + 1] : n;
    const uint32_t count = (uint32_t)(end - start);
    const uint64_t code = sigk_key_code(keys[start]);

    // Boyer-Moore vote over func_index (replaces func_count, tcc:203)
    uint32_t cand = 0, votes = 0;
    for (uint64_t j = start; j < end; ++j) {
        const uint32_t f = __ldg(meta.func + vals[j]);
        if (votes == 0) { cand = f; votes = 1; }
        else if (f == cand) ++votes;
        else --votes;
    }
    uint32_t best_count = 0;
    if (count == 1) best_count = 1;
    else for (uint64_t j = start; j < end; ++j) best_count += (__ldg(meta.func + vals[j]) == cand);

    bool keep = 2ull * best_count > count;
    if (keep) {
        const float thresh = __fmul_rn(__int2float_rn((int)count), 0.8f);       // tcc:250
        keep = !(__int2float_rn((int)best_count) < thresh);                     // tcc:254
    }
    if (!keep) { seg_rows[s] = make_uint4(0, 0, 0, 0); return; }

    // avg_from_end = sorted(offsets of ALL items)[count/2]   (tcc:273, :281-282)
    uint32_t avg;
    if (count == 1) avg = sigk_key_offset(keys[start]);
    else {
        const uint32_t r = count / 2;
        uint32_t lo = 0, hi = 0xFFFFu;
        while (lo < hi) {                       // smallest v with #(offset <= v) > r
            const uint32_t mid = (lo + hi) >> 1;
            uint32_t c = 0;
            for (uint64_t j = start; j < end; ++j) c += (sigk_key_offset(keys[j]) <= mid);
            if (c > r) hi = mid; else lo = mid + 1;
        }
        avg = lo;
    }

    // length statistics over the best function's items, newest first; every item marks its sequence
    LengthAcc acc;
    uint32_t S = 0;
    for (uint64_t j = end; j-- > start;) {
        const uint32_t ord = vals[j];
        if (__ldg(meta.func + ord) == cand) {
            const uint32_t len = __ldg(meta.len + ord);
            if (order_stats) acc.push(len);
            else S = (S + len) & 0xFFFFu;
        }
        const uint32_t sid = __ldg(meta.seq_id + ord);      // seqs_with_a_signature.insert, tcc:274
        const uint32_t bit = 1u << (sid & 31u);
        if (!(seq_bitmap[sid >> 5] & bit)) atomicOr(seq_bitmap + (sid >> 5), bit);
    }
    if (order_stats) S = acc.S;
    const uint32_t mean = S / best_count;                   // u16((double)S / n): exact
    const uint32_t median = order_stats ? u16_from_double(acc.q2) : 0u;
    const uint32_t var = order_stats ? u16_from_double(acc.var) : 0u;

    atomicAdd(distinct_functions + cand, 1u);               // tcc:286
    seg_rows[s] = make_uint4((uint32_t)code, (uint32_t)(code >> 32) | (avg << 11) | (1u << 31),
                             cand | (mean << 16), median | (var << 16));
}

constexpr int CMP_THREADS = 256;
constexpr int CMP_ITEMS = CMP_TILE / CMP_THREADS;      // 8
constexpr int CMP_WARPS = CMP_THREADS / 32;

__global__ void __launch_bounds__(CMP_THREADS)
compact_rows_kernel(const uint4 *__restrict__ seg_rows, const uint64_t *__restrict__ n_seg_ptr, KeptColumns out,
                    uint64_t *__restrict__ scan_state, uint32_t *__restrict__ ticket, uint64_t *__restrict__ n_kept_out) {
    __shared__ uint32_t s_scan[CMP_WARPS + 2];
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_base;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint64_t n_seg = *n_seg_ptr;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t tile_start = (uint64_t)tile * CMP_TILE;
    if (tile_start >= n_seg) return;

    uint4 row[CMP_ITEMS];
    unsigned ball[CMP_ITEMS];
    const uint64_t wbase = tile_start + (uint64_t)warp * (CMP_ITEMS * 32);
#pragma unroll
    for (int i = 0; i < CMP_ITEMS; ++i) {
        const uint64_t s = wbase + i * 32 + lane;
        row[i] = s < n_seg ? seg_rows[s] : make_uint4(0, 0, 0, 0);
        ball[i] = __ballot_sync(0xffffffffu, (row[i].y >> 31) != 0);
    }
    uint32_t total;
    uint32_t run = striped_flag_scan<CMP_THREADS, CMP_ITEMS>(ball, s_scan, &total);
    if (tid == 0) {
        const uint64_t base = chained_scan_exclusive(scan_state, tile, total);
        s_base = base;
        if (tile_start + CMP_TILE >= n_seg) *n_kept_out = base + total;
    }
    __syncthreads();
    const uint64_t base = s_base;
#pragma unroll
    for (int i = 0; i < CMP_ITEMS; ++i) {
        if ((ball[i] >> lane) & 1u) {
            const uint64_t o = base + run + __popc(ball[i] & ((1u << lane) - 1u));
            const uint64_t code = (uint64_t)row[i].x | ((uint64_t)(row[i].y & 0x7FFu) << 32);
            out.kmer[o] = sigk_code_to_ascii(code);
            out.avg_from_end[o] = (uint16_t)((row[i].y >> 11) & 0xFFFFu);
            out.function_index[o] = (uint16_t)(row[i].z & 0xFFFFu);
            out.mean[o] = (uint16_t)(row[i].z >> 16);
            out.median[o] = (uint16_t)(row[i].w & 0xFFFFu);
            out.var[o] = (uint16_t)(row[i].w >> 16);
        }
        run += __popc(ball[i]);
    }
}

__global__ void popcount_kernel(const uint32_t *__restrict__ bitmap, uint64_t n_words, uint64_t *__restrict__ out) {
    uint64_t c = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (uint64_t)gridDim.x * blockDim.x)
        c += __popc(bitmap[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31u) == 0 && c) atomicAdd(reinterpret_cast<unsigned long long *>(out), (unsigned long long)c);
}

}  // namespace

cudaError_t launch_segment_heads(const uint64_t *keys, const uint64_t *n_ptr, uint64_t capacity, uint32_t *seg_start,
                                 uint64_t *scan_state, uint32_t *ticket, uint64_t *n_seg_out, cudaStream_t stream) {
    if (capacity == 0) return cudaSuccess;
    segment_heads_kernel<<<(unsigned)seg_tiles(capacity), SEG_THREADS, 0, stream>>>(keys, n_ptr, seg_start, scan_state, ticket, n_seg_out);
    return cudaGetLastError();
}

cudaError_t launch_segment_process(const uint64_t *keys, const uint32_t *vals, const uint64_t *n_ptr,
                                   const uint32_t *seg_start, const uint64_t *n_seg_ptr, uint64_t capacity,
                                   ProteinMeta meta, int order_stats, uint4 *seg_rows, uint32_t *seq_bitmap,
                                   uint32_t *distinct_functions, cudaStream_t stream) {
    if (capacity == 0) return cudaSuccess;
    const uint64_t blocks = (capacity + 255) / 256;
    segment_process_kernel<<<(unsigned)blocks, 256, 0, stream>>>(keys, vals, n_ptr, seg_start, n_seg_ptr, meta, order_stats,
                                                                 seg_rows, seq_bitmap, distinct_functions);
    return cudaGetLastError();
}

cudaError_t launch_compact_rows(const uint4 *seg_rows, const uint64_t *n_seg_ptr, uint64_t capacity, KeptColumns out,
                                uint64_t *scan_state, uint32_t *ticket, uint64_t *n_kept_out, cudaStream_t stream) {
    if (capacity == 0) return cudaSuccess;
    compact_rows_kernel<<<(unsigned)cmp_tiles(capacity), CMP_THREADS, 0, stream>>>(seg_rows, n_seg_ptr, out, scan_state, ticket, n_kept_out);
    return cudaGetLastError();
}

cudaError_t launch_popcount(const uint32_t *bitmap, uint64_t n_words, uint64_t *out, cudaStream_t stream) {
    if (n_words == 0) return cudaSuccess;
    uint64_t blocks = (n_words + 255) / 256;
    if (blocks > 1024) blocks = 1024;
    popcount_kernel<<<(unsigned)blocks, 256, 0, stream>>>(bitmap, n_words, out);
    return cudaGetLastError();
}

}  // namespace sigk
