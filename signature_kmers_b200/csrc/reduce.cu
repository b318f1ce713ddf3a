// reduce.cu — stages 3 and 4: run-length the sorted records into k-mer groups,
// reduce each group (function tally, offset median, length sum), apply the
// reference's keep/reject rule and compact the kept rows into the table; a
// follow-up kernel fills the order-dependent median/var columns.
//
// Replaces SignatureBuilder<K>::process_kmers / process_kmer_set
// (reference src/signature_build.tcc:183-293):
//   key-change walk       :194-207         -> head flags by neighbour compare
//   tally + arg-max       :203, :228-248   -> bit-sliced majority vote (below)
//   80 % rule             :250-257         -> same float32 ops
//   offsets median        :273, :281-282   -> rank n/2 by bit-sliced radix select
//   mean                  :268-271, :277   -> 16-bit wrapping sum / best_count
//   median / var          :262-264, :278-279 -> order_stats_kernel (length_acc.cuh)
//   kept row              :288             -> StoredKmerData columns, k-mer order
//   statistics            :274, :285-286   -> per-protein kept-occurrence test, per-function counts
//
// Majority vote instead of a tally: a group is kept only if best_count >=
// 0.8*count, so a kept group's best function is a strict majority.  For a strict
// majority element every bit of its index is the majority value of that bit, so
// 16 ballot+popc steps give the only possible candidate; counting it decides.
// If it is not a strict majority no function is, the true best_count is
// <= count/2 and the float test rejects for every count — ties never reach a
// kept row, so the std::map walk's "lowest index wins" cannot matter.
//
// Work mapping: head_tile_kernel (count pass, emit pass) run-lengths the records in independent
// 2048-record tiles and finishes single-record groups; group_reduce_kernel packs the groups of
// 2..32 records 32 records to a warp (lane = record, every group of the window reduced at once
// with ballots and segmented shuffles) and walks longer groups with a whole warp; every group
// writes the row slot of its index (a tombstone when rejected); order_stats_kernel patches median
// and var into the rows that need the ordered walk; squeeze_rows_kernel drops the tombstones and
// expands the rows into the table columns; function_histogram_kernel tallies distinct_functions.
// No kernel of this stage waits for another tile: bases come from two one-block scans of per-tile
// counters (tile_scan_kernel, tombstone_scan_kernel).
//
// HBM traffic per record: 8 B (count pass) + 12 B read (+ an 8-byte L2 gather of the protein's
// meta); per group a 16-byte row written and read once; per kept row 18 B written.
#include "kernels.h"
#include "sigk_common.cuh"
#include "length_acc.cuh"

#include <algorithm>

namespace sigk {

namespace {

constexpr unsigned FULL = 0xffffffffu;

SIGK_D unsigned mask_lt(unsigned lane) { return (1u << lane) - 1u; }
SIGK_D unsigned mask_le(unsigned lane) { return (2u << lane) - 1u; }      // lane 31 -> 0xffffffff (shift wraps to 0, -1)
SIGK_D unsigned mask_range(unsigned lo, unsigned hi) {                     // bits [lo, hi), hi <= 32
    const unsigned upper = hi >= 32 ? FULL : ((1u << hi) - 1u);
    return upper & ~((1u << lo) - 1u);
}

// The per-protein table comes in two widths (kernels.h): 8 bytes {length, function}, or 4 bytes
// length | function << 16 when every protein of the job is shorter than 65 535 residues (half the
// footprint: what decides whether the gathers stay in L2 once several ranks' proteins are in it).
// (c_meta_shift: SIGK_TEST_META_SPREAD=k spreads the entries 2^k apart, so that a one-GPU run can be given the cache
// footprint of the job-wide table of a many-GPU run; 0 in production)
__constant__ uint32_t c_meta_shift;
template <typename MetaT> SIGK_D ProtMeta load_meta(const MetaT *__restrict__ meta, uint32_t ordinal);
template <> SIGK_D ProtMeta load_meta<ProtMeta>(const ProtMeta *__restrict__ meta, uint32_t ordinal) { return ld_keep_u32x2(meta + ((size_t)ordinal << c_meta_shift)); }
template <> SIGK_D ProtMeta load_meta<uint32_t>(const uint32_t *__restrict__ meta, uint32_t ordinal) {
    const uint32_t v = ld_keep_u32(meta + ((size_t)ordinal << c_meta_shift));
    return make_uint2(v & 0xFFFFu, v >> 16);
}

// the sorted records stream through this stage once per pass
#ifndef SIGK_RED_EVICT_FIRST
#define SIGK_RED_EVICT_FIRST 0
#endif
#if SIGK_RED_EVICT_FIRST
SIGK_D uint64_t rec_u64(const uint64_t *p) { return ld_once_u64(p); }
SIGK_D uint32_t rec_u32(const uint32_t *p) { return ld_once_u32(p); }
SIGK_D ulonglong2 rec_u64x2(const ulonglong2 *p) { return ld_once_u64x2(p); }
SIGK_D uint4 rec_u128(const uint4 *p) { return ld_once_u128(p); }
#else
SIGK_D uint64_t rec_u64(const uint64_t *p) { return __ldg(p); }
SIGK_D uint32_t rec_u32(const uint32_t *p) { return __ldg(p); }
SIGK_D ulonglong2 rec_u64x2(const ulonglong2 *p) { return __ldg(p); }
SIGK_D uint4 rec_u128(const uint4 *p) { return __ldg(p); }
#endif

SIGK_D bool keep_rule(uint32_t best_count, uint32_t count) {
    if (2ull * best_count <= count) return false;                          // no strict majority (see header)
    // reject iff (float)best_count < float(count) * 0.8f (tcc:250-257).  0.8*count is at least 0.2
    // away from any integer it does not equal and the float product is off by < count * 6e-8, so
    // below 2^20 the float test is exactly 5*best >= 4*count; above, run the float ops themselves.
    if (count < (1u << 20)) return 5u * best_count >= 4u * count;
    const float thresh = __fmul_rn(__int2float_rn((int)count), 0.8f);
    return !(__int2float_rn((int)best_count) < thresh);
}

// seqs_with_a_signature (tcc:274) without touching a bitmap per record: a protein has a signature iff
// at least one of its occurrences lies in a kept group.  encode counts each protein's occurrences,
// the reduce kernels count the (rare) occurrences that fall into rejected groups, and
// signature_flags_kernel compares the two per protein.
// (c_rej_shift: SIGK_TEST_REJ_SPREAD=k spreads the counters 2^k apart — the cache footprint of a many-GPU job's counter array on one GPU; 0 in production)
__constant__ uint32_t c_rej_shift;
SIGK_D void count_rejected(uint32_t *prot_rejected, uint32_t ordinal) { atomicAdd(prot_rejected + ((size_t)ordinal << c_rej_shift), 1u); }

struct SegResult {
    bool keep;
    bool closed;            // all best lengths equal and the 16-bit sum did not wrap: median = that length, var = 0
    uint32_t func, best_count, avg, mean, len0;
};

// A group of n > 32 records starting at `start`, reduced by the whole warp in strides of 32:
// one walk that counts the first record's function, sums its lengths and finds the offset bits that
// vary (a vote and a recount follow only if some record has another function), then one walk per
// varying offset bit for the radix select (inside a family the offsets rarely differ in more than a
// few bits; none when they are all equal).
template <typename MetaT>
SIGK_D SegResult reduce_long_segment(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals,
                                     const MetaT *__restrict__ meta, uint64_t start, uint32_t n, uint32_t *prot_rejected) {
    const unsigned lane = threadIdx.x & 31u;
    SegResult r{false, false, 0, 0, 0, 0, 0};
    // walk 1, for the usual case of a group with one function (a k-mer of one family): take the first record's
    // function as the candidate; count it, sum its lengths (mod 65536), OR the offset differences, and notice any
    // record that disagrees.  Only then is a vote needed.
    const uint32_t off0 = sigk_key_offset(keys[start]);
    uint32_t cand = load_meta(meta, vals[start]).y;
    uint32_t best = 0, S = 0, vary = 0, lmin = 0xFFFFFFFFu, lmax = 0;
    bool mixed = false;
    for (uint32_t base = 0; base < n; base += 32) {
        const uint32_t j = base + lane;
        if (j < n) {
            const ProtMeta m = load_meta(meta, vals[start + j]);
            if (m.y == cand) { ++best; S += m.x; lmin = min(lmin, m.x); lmax = max(lmax, m.x); }
            else mixed = true;
            vary |= sigk_key_offset(keys[start + j]) ^ off0;
        }
    }
    if (__any_sync(FULL, mixed)) {
        // walk 2: bit-sliced majority vote over func_index; lane b owns bit b
        uint32_t ones = 0;
        for (uint32_t base = 0; base < n; base += 32) {
            const uint32_t j = base + lane;
            const bool act = j < n;
            const uint32_t f = act ? load_meta(meta, vals[start + j]).y : 0u;
#pragma unroll
            for (int b = 0; b < 16; ++b) {
                const unsigned bal = __ballot_sync(FULL, act && ((f >> b) & 1u));
                if ((int)lane == b) ones += __popc(bal);
            }
        }
        const uint32_t voted = __ballot_sync(FULL, lane < 16 && 2ull * ones > n) & 0xFFFFu;
        if (voted != cand) {
            // walk 3: the first record's function was not the majority: count the voted one instead
            cand = voted;
            best = 0; S = 0; lmin = 0xFFFFFFFFu; lmax = 0;
            for (uint32_t base = 0; base < n; base += 32) {
                const uint32_t j = base + lane;
                if (j < n) {
                    const ProtMeta m = load_meta(meta, vals[start + j]);
                    if (m.y == cand) { ++best; S += m.x; lmin = min(lmin, m.x); lmax = max(lmax, m.x); }
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { best += __shfl_xor_sync(FULL, best, o); S += __shfl_xor_sync(FULL, S, o); }
    vary = __reduce_or_sync(FULL, vary);
    lmin = __reduce_min_sync(FULL, lmin);
    lmax = __reduce_max_sync(FULL, lmax);
    r.func = cand;
    r.best_count = best;
    r.keep = keep_rule(best, n);
    if (!r.keep) {
        for (uint32_t j = lane; j < n; j += 32) count_rejected(prot_rejected, vals[start + j]);
        return r;
    }
    r.mean = (S & 0xFFFFu) / best;
    r.len0 = lmin;
    r.closed = lmin == lmax && (uint64_t)best * lmin < 65536ull;
    // offset of rank n/2 by radix select over the varying bits, most significant first: one walk per bit
    uint32_t rank = n / 2, prefix = off0 & ~vary, pmask = ~vary & 0xFFFFu;
    while (vary) {
        const int b = 31 - __clz(vary);
        uint32_t zeros = 0;
        for (uint32_t j = lane; j < n; j += 32) {
            const uint32_t off = sigk_key_offset(keys[start + j]);
            zeros += ((off & pmask) == prefix && !((off >> b) & 1u)) ? 1u : 0u;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) zeros += __shfl_xor_sync(FULL, zeros, o);
        if (rank >= zeros) { rank -= zeros; prefix |= 1u << b; }
        pmask |= 1u << b;
        vary &= ~(1u << b);
    }
    r.avg = prefix;
    return r;
}

// ---- stage 3a: run-length ---------------------------------------------------------------------
// One tile of RED_BATCH sorted records per block, HS_ITEMS consecutive records per thread: head
// flags by neighbour compare; a head's group ends at the next head, which a suffix-min over the block
// gives to every thread.  Every group, kept or not, owns the row slot of its index, so nothing
// downstream waits for a keep decision; rejected groups leave a tombstone (function_index 0xFFFF)
// that squeeze_rows_kernel removes.
//
// The kernel runs twice.  The COUNT pass reads the keys only and leaves per tile {heads, groups of
// 2..32 records}; tile_scan_kernel turns the counts into row and list bases; the EMIT pass recomputes
// the flags and writes.  Reading the keys a second time (8 of the 153 algorithmic bytes per record)
// buys tiles that are independent of each other: the single-pass version with a chained scan over
// the tiles spent most of its time waiting for predecessors (profiles/r1_ncu_full_reduce_v9.txt).
// Single-record groups (92 % of the groups of the 2 M-protein set, always kept: 1 >= 0.8) are finished
// in the EMIT pass; groups of 2..32 records go to `groups` (in position order), longer ones to
// `long_groups`.  The last group of a tile may run into the following tiles: the COUNT pass leaves
// its position in tile_open and the position of every tile's first head in tile_first, and
// resolve_open_kernel closes those groups afterwards (no look-ahead, so a group of millions of
// records costs nothing extra).
//
// rows[g] (uint4): x = code[31:0]; y = code[42:32] | avg_from_end << 11;
//                  z = function_index | mean << 16; w = median | var << 16
constexpr int RED_THREADS = 128;
constexpr int HS_THREADS = 256;
constexpr int HS_ITEMS = 8;
constexpr int HS_WARPS = HS_THREADS / 32;
static_assert(HS_THREADS * HS_ITEMS == RED_BATCH, "one tile per scan entry");
constexpr int WORK_BLOCK = 64;          // list slots a warp reserves at a time (order-statistics work)
constexpr int GR_CHUNK = 256;           // group descriptors a warp takes per fetch
#ifndef SIGK_ORD_LONG
#define SIGK_ORD_LONG 32768
#endif
constexpr uint32_t ORD_LONG = SIGK_ORD_LONG;    // groups above this are walked by a whole warp (the tail); below, a lane each
constexpr uint32_t HS_NONE = 0xFFFFFFFFu;
constexpr int SCAN_THREADS = 1024;      // the single block that scans the per-tile counters
constexpr int SQ_THREADS = 256;
constexpr int SQ_ITEMS = 8;
constexpr int SQ_TILE = SQ_THREADS * SQ_ITEMS;
constexpr int SQ_TILE_SHIFT = 11;
constexpr int SQ_WARPS = SQ_THREADS / 32;
static_assert(SQ_TILE == 1 << SQ_TILE_SHIFT, "squeeze tile");

struct WorkCursor { uint32_t base, free; };

// Slots are reserved in blocks so that the shared counter sees one atomic per 64 entries; the unused
// tail of a block is filled with count-0 entries (consumers skip them).  Inside a block the used
// entries are a prefix.
SIGK_D void work_reserve(WorkCursor &wc, uint32_t need, OrderWork *__restrict__ work, uint32_t *__restrict__ n_work) {
    const unsigned lane = threadIdx.x & 31u;
    if (wc.free >= need) return;
    for (uint32_t i = lane; i < wc.free; i += 32) work[wc.base + i] = OrderWork{0u, 0u, 0u};
    uint32_t b = 0;
    if (lane == 0) b = atomicAdd(n_work, (uint32_t)WORK_BLOCK);
    wc.base = __shfl_sync(FULL, b, 0);
    wc.free = WORK_BLOCK;
}
SIGK_D void work_flush(WorkCursor &wc, OrderWork *__restrict__ work) {
    for (uint32_t i = threadIdx.x & 31u; i < wc.free; i += 32) work[wc.base + i] = OrderWork{0u, 0u, 0u};
}

SIGK_D uint4 singleton_row(uint64_t key, const ProtMeta m) {
    const uint64_t code = sigk_key_code(key);
    // one item: avg_from_end = its offset, best = its function, sum = its length mod 65536, median = var = 0
    return make_uint4((uint32_t)code, (uint32_t)(code >> 32) | (sigk_key_offset(key) << 11), m.y | ((m.x & 0xFFFFu) << 16), 0u);
}

#ifndef SIGK_HS_MIN_BLOCKS
#define SIGK_HS_MIN_BLOCKS 4
#endif
template <bool EMIT, typename MetaT>
__global__ void __launch_bounds__(HS_THREADS, SIGK_HS_MIN_BLOCKS)
head_tile_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, const uint64_t *__restrict__ n_ptr,
                 const MetaT *__restrict__ meta, uint4 *__restrict__ rows, OrderWork *__restrict__ groups,
                 OrderWork *__restrict__ long_groups, uint32_t *__restrict__ n_long, ReduceScratch sx) {
    __shared__ uint32_t s_scan[HS_WARPS + 2];
    __shared__ uint64_t s_last[HS_WARPS];
    __shared__ uint32_t s_wfirst[HS_WARPS];
    __shared__ uint4 s_rows[EMIT ? RED_BATCH : 1];
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint64_t n = *n_ptr;
    constexpr uint64_t NO_CODE = ~0ull;                  // codes are < 2^43
    const uint32_t tile = blockIdx.x;
    const uint64_t tile_start = (uint64_t)tile * RED_BATCH;
    if (tile_start >= n) return;
    const uint32_t tile_n = (uint32_t)(n - tile_start < (uint64_t)RED_BATCH ? n - tile_start : (uint64_t)RED_BATCH);
    const bool last_tile = tile_start + tile_n == n;
    const uint32_t t0 = tid * HS_ITEMS;
    const uint64_t base = EMIT ? __ldg(sx.tile_base + tile) : 0ull;      // rows | list slots << 32

    uint64_t before = NO_CODE;                           // code of the record in front of the tile
    if (tid == 0 && tile_start) before = sigk_key_code(__ldg(keys + tile_start - 1));
    uint64_t k[HS_ITEMS];
    uint32_t v[HS_ITEMS];
    if (tile_n == RED_BATCH) {
        const ulonglong2 *kp = reinterpret_cast<const ulonglong2 *>(keys + tile_start + t0);
#pragma unroll
        for (int i = 0; i < HS_ITEMS / 2; ++i) { const ulonglong2 x = rec_u64x2(kp + i); k[2 * i] = x.x; k[2 * i + 1] = x.y; }
        if (EMIT) {
            const uint4 *vp = reinterpret_cast<const uint4 *>(vals + tile_start + t0);
#pragma unroll
            for (int i = 0; i < HS_ITEMS / 4; ++i) {
                const uint4 x = rec_u128(vp + i);
                v[4 * i] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < HS_ITEMS; ++i) {
            const bool ok = t0 + i < tile_n;
            k[i] = ok ? rec_u64(keys + tile_start + t0 + i) : ~0ull;
            if (EMIT) v[i] = ok ? rec_u32(vals + tile_start + t0 + i) : 0u;
        }
    }
    if (lane == 31) s_last[warp] = sigk_key_code(k[HS_ITEMS - 1]);
    __syncthreads();
    uint64_t prev = __shfl_up_sync(FULL, sigk_key_code(k[HS_ITEMS - 1]), 1);
    if (lane == 0) prev = warp ? s_last[warp - 1] : before;
    uint32_t hmask = 0;
#pragma unroll
    for (int i = 0; i < HS_ITEMS; ++i) {
        const uint64_t c = sigk_key_code(k[i]);
        if (t0 + i < tile_n && c != prev) hmask |= 1u << i;             // kmer != cur, tcc:194
        prev = c;
    }
    const uint32_t hc = __popc(hmask);
    // meta of every head's protein, requested now so that the gathers fly during the scan below (only the
    // single-record groups, 92 % of them, use it)
    ProtMeta m[HS_ITEMS];
    if (EMIT) {
#pragma unroll
        for (int i = 0; i < HS_ITEMS; ++i) m[i] = ((hmask >> i) & 1u) ? load_meta(meta, v[i]) : make_uint2(0, 0);
    }

    // next head after this thread's records (local position), HS_NONE if the tile has none
    uint32_t fh = hmask ? t0 + (uint32_t)__ffs(hmask) - 1u : HS_NONE;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_down_sync(FULL, fh, o);
        if (lane + o < 32) fh = min(fh, y);
    }
    uint32_t nh = __shfl_down_sync(FULL, fh, 1);
    if (lane == 31) nh = HS_NONE;
    if (lane == 0) s_wfirst[warp] = fh;
    __syncthreads();
#pragma unroll
    for (int w = 1; w < HS_WARPS; ++w) nh = min(nh, (int)warp + w < HS_WARPS ? s_wfirst[(warp + w) & (HS_WARPS - 1)] : HS_NONE);
    if (nh == HS_NONE && last_tile) nh = tile_n;         // the last group of the job ends with the records

    // length of every group that starts here (0 = runs past the tile)
    uint32_t cnt[HS_ITEMS];
    uint32_t mc = 0;
#pragma unroll
    for (int i = 0; i < HS_ITEMS; ++i) {
        cnt[i] = 0;
        if ((hmask >> i) & 1u) {
            const uint32_t above = hmask >> (i + 1);
            const uint32_t next = above ? t0 + (uint32_t)i + (uint32_t)__ffs(above) : nh;
            cnt[i] = next == HS_NONE ? 0u : next - (t0 + (uint32_t)i);
            mc += (cnt[i] >= 2 && cnt[i] <= 32) ? 1u : 0u;
            if (!EMIT && cnt[i] == 0) sx.tile_open[tile] = (uint32_t)(tile_start + t0 + i) + 1u;
        }
    }
    // one scan for both counters: heads in the low half, listed groups in the high half (a tile has at
    // most 2048 of either)
    uint32_t total;
    const uint32_t excl = block_exclusive_scan<HS_THREADS>(hc | (mc << 16), s_scan, &total);
    if (!EMIT) {
        if (tid == 0) {
            sx.tile_counts[tile] = (uint64_t)(total & 0xFFFFu) | ((uint64_t)(total >> 16) << 32);
            uint32_t first = HS_NONE;
#pragma unroll
            for (int w = 0; w < HS_WARPS; ++w) first = min(first, s_wfirst[w]);
            if (first != HS_NONE) sx.tile_first[tile] = (uint32_t)(tile_start + first) + 1u;
        }
        return;
    }
    // The tile's rows are contiguous (one per head): they are assembled in shared memory and leave as whole
    // 512-byte warp stores.  Slots of groups that are not finished here get a tombstone for now; the kernel
    // that finishes them (group_reduce_kernel, resolve_open_kernel) runs later and overwrites it.
    const uint64_t g0 = base & 0xFFFFFFFFull;
    uint32_t gl = excl & 0xFFFFu;                        // row slot inside the tile
    uint32_t gslot = (uint32_t)(base >> 32) + (excl >> 16);
#pragma unroll
    for (int i = 0; i < HS_ITEMS; ++i) {
        if ((hmask >> i) & 1u) {
            const uint32_t p = (uint32_t)(tile_start + t0 + i);
            if (cnt[i] == 1) {
                s_rows[gl] = singleton_row(k[i], m[i]);
            } else {
                s_rows[gl] = make_uint4(0u, 0u, 0xFFFFu, 0u);
                if (cnt[i] == 0) {
                    // closed by resolve_open_kernel
                } else if (cnt[i] <= 32) {
                    groups[gslot++] = OrderWork{(uint32_t)(g0 + gl), p, cnt[i]};
                } else {
                    long_groups[atomicAdd(n_long, 1u)] = OrderWork{(uint32_t)(g0 + gl), p, cnt[i]};
                }
            }
            ++gl;
        }
    }
    __syncthreads();
    const uint32_t n_rows = total & 0xFFFFu;
    for (uint32_t j = tid; j < n_rows; j += HS_THREADS) rows[g0 + j] = s_rows[j];
}

// Exclusive scan of `count` per-tile counters by one block, written for coalesced access: warp w owns the
// contiguous chunk w of the tiles and walks it 32 entries at a time (once for the chunk total, once to
// write the prefixes).  Returns the grand total in every thread.
template <typename T>
SIGK_D T block_scan_tiles(const T *__restrict__ in, T *__restrict__ out, uint64_t count, T *s_part /* 33 */) {
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint64_t per = ((count + SCAN_THREADS / 32 - 1) / (SCAN_THREADS / 32) + 31) & ~31ull;    // chunk, a multiple of 32
    const uint64_t lo = min(count, (uint64_t)warp * per), hi = min(count, lo + per);
    T sum = 0;
    for (uint64_t t = lo + lane; t < hi; t += 32) sum += in[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
    if (lane == 0) s_part[warp] = sum;
    __syncthreads();
    if (warp == 0) {
        const T w = s_part[lane];
        T wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const T y = __shfl_up_sync(FULL, wi, o);
            if (lane >= (unsigned)o) wi += y;
        }
        s_part[lane] = wi - w;
        if (lane == 31) s_part[32] = wi;
    }
    __syncthreads();
    T run = s_part[warp];
    for (uint64_t t0 = lo; t0 < hi; t0 += 32) {
        const uint64_t t = t0 + lane;
        const T x = t < hi ? in[t] : (T)0;
        T incl = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const T y = __shfl_up_sync(FULL, incl, o);
            if (lane >= (unsigned)o) incl += y;
        }
        if (t < hi) out[t] = run + incl - x;
        run += __shfl_sync(FULL, incl, 31);
    }
    return s_part[32];
}

// tile_base[t] = {heads, listed groups} before tile t; the totals are the number of groups and of list entries.
__global__ void __launch_bounds__(SCAN_THREADS)
tile_scan_kernel(const uint64_t *__restrict__ n_ptr, ReduceScratch sx, uint64_t *__restrict__ n_seg_out, uint32_t *__restrict__ n_groups) {
    __shared__ uint64_t s_part[SCAN_THREADS / 32 + 1];
    const uint64_t n = *n_ptr;
    const uint64_t total = block_scan_tiles<uint64_t>(sx.tile_counts, sx.tile_base, (n + RED_BATCH - 1) / RED_BATCH, s_part);
    if (threadIdx.x == 0) {
        *n_seg_out = total & 0xFFFFFFFFull;
        *n_groups = (uint32_t)(total >> 32);
    }
}

// Groups that ran past the end of their tile (always the tile's last head): the next head is the first
// head of the next tile that has one (tile_first, 0 = none), or the end of the records.
template <typename MetaT>
__global__ void resolve_open_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals,
                                    const uint64_t *__restrict__ n_ptr, const MetaT *__restrict__ meta, ReduceScratch sx,
                                    uint4 *__restrict__ rows, OrderWork *__restrict__ groups, uint32_t *__restrict__ n_groups,
                                    OrderWork *__restrict__ long_groups, uint32_t *__restrict__ n_long) {
    const uint64_t n = *n_ptr;
    const uint64_t n_tiles = (n + RED_BATCH - 1) / RED_BATCH;
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const uint32_t o = sx.tile_open[t];
    if (!o) return;
    const uint32_t start = o - 1u;
    const uint32_t g = (uint32_t)sx.tile_base[t] + (uint32_t)sx.tile_counts[t] - 1u;
    uint64_t end = n;
    for (uint64_t u = t + 1; u < n_tiles; ++u) {
        const uint32_t f = __ldg(sx.tile_first + u);
        if (f) { end = f - 1u; break; }
    }
    const uint32_t cnt = (uint32_t)(end - start);
    if (cnt == 1) {
        const ProtMeta m = load_meta(meta, __ldg(vals + start));
        rows[g] = singleton_row(__ldg(keys + start), m);
    } else if (cnt <= 32) {
        groups[atomicAdd(n_groups, 1u)] = OrderWork{g, start, cnt};
    } else {
        long_groups[atomicAdd(n_long, 1u)] = OrderWork{g, start, cnt};
    }
}

// ---- stage 3b: groups of 2..32 records, packed 32 records to a warp ---------------------------
// A warp takes 32 descriptors, lays their records side by side (lane = record, whole groups only)
// and reduces all groups of the window at once with ballots and segmented shuffles.
template <typename MetaT>
__global__ void __launch_bounds__(RED_THREADS)
group_reduce_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, const MetaT *__restrict__ meta,
                    const OrderWork *__restrict__ groups, const uint32_t *__restrict__ n_groups, uint32_t *__restrict__ next_group,
                    const OrderWork *__restrict__ long_groups, const uint32_t *__restrict__ n_long, uint32_t *__restrict__ next_long,
                    uint4 *__restrict__ rows, OrderWork *__restrict__ work, uint32_t *__restrict__ n_work,
                    OrderWork *__restrict__ work_long, uint32_t *__restrict__ n_work_long, OrderWork *__restrict__ giant,
                    uint32_t *__restrict__ n_giant, uint32_t *__restrict__ prot_rejected,
                    uint32_t *__restrict__ rej_tile, int order_stats) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t total = *n_groups;
    WorkCursor wc{0u, 0u};

    // ---- groups of more than 32 records first (they are the tail), one warp each
    {
        const uint32_t total_long = *n_long;
        for (;;) {
            uint32_t i = 0;
            if (lane == 0) i = atomicAdd(next_long, 1u);
            i = __shfl_sync(FULL, i, 0);
            if (i >= total_long) break;
            const OrderWork d = long_groups[i];
            if (d.count > REDUCE_GIANT) {           // a warp would walk it for milliseconds: giant_reduce_kernel, a CTA per group
                if (lane == 0) giant[atomicAdd(n_giant, 1u)] = d;
                continue;
            }
            const SegResult r = reduce_long_segment(keys, vals, meta, d.start, d.count, prot_rejected);
            const bool walk = r.keep && order_stats && !r.closed;   // best_count >= 27 here
            const bool walk_long = walk && d.count > ORD_LONG;      // whole-warp walk (order_stats_long_kernel)
            if (walk && !walk_long) work_reserve(wc, 1u, work, n_work);
            if (lane == 0) {
                if (r.keep) {
                    const uint64_t code = sigk_key_code(keys[d.start]);
                    rows[d.row] = make_uint4((uint32_t)code, (uint32_t)(code >> 32) | (r.avg << 11), r.func | (r.mean << 16),
                                             (order_stats && r.closed) ? r.len0 : 0u);
                    if (walk_long) work_long[atomicAdd(n_work_long, 1u)] = d;
                    else if (walk) work[wc.base] = d;
                } else {
                    rows[d.row] = make_uint4(0u, 0u, 0xFFFFu, 0u);
                    atomicAdd(rej_tile + (d.row >> SQ_TILE_SHIFT), 1u);         // the squeeze's per-tile tombstone count
                }
            }
            if (walk && !walk_long) { wc.base += 1; wc.free -= 1; }
        }
    }

    // ---- then the packed groups, GR_CHUNK descriptors per fetch
    for (;;) {
      uint32_t c0 = 0;
      if (lane == 0) c0 = atomicAdd(next_group, (uint32_t)GR_CHUNK);
      c0 = __shfl_sync(FULL, c0, 0);
      if (c0 >= total) break;
      for (uint64_t d0 = c0; d0 < (uint64_t)c0 + GR_CHUNK && d0 < total; d0 += 32) {
        const uint64_t di = d0 + lane;
        OrderWork d = di < total ? groups[di] : OrderWork{0u, 0u, 0u};
        // used descriptors are a prefix of the 32 (work_reserve pads block tails)
        uint32_t incl = d.count;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL, incl, o);
            if (lane >= (unsigned)o) incl += y;
        }
        const uint32_t excl = incl - d.count;
        const uint32_t tot = __shfl_sync(FULL, incl, 31);
        uint32_t base = 0, j0 = 0;
        while (base < tot) {
            // ---- one window: descriptors j0.. whose records fit into 32 lanes
            const bool fits = lane >= j0 && d.count && incl - base <= 32u;
            const unsigned F = __ballot_sync(FULL, fits);
            const uint32_t nd = __popc(F);
            const uint32_t wlen = __shfl_sync(FULL, incl, j0 + nd - 1) - base;
            const unsigned H = __reduce_or_sync(FULL, fits ? (1u << (excl - base)) : 0u);
            const bool act = lane < wlen;
            const unsigned s_lane = 31u - (unsigned)__clz(H & mask_le(lane));
            const uint32_t dj = j0 + __popc(H & mask_le(lane)) - 1u;        // my group's descriptor lane
            const uint32_t cnt = __shfl_sync(FULL, d.count, dj);
            const uint32_t gstart = __shfl_sync(FULL, d.start, dj);
            const uint32_t grow = __shfl_sync(FULL, d.row, dj);
            const unsigned e_lane = s_lane + cnt;
            const unsigned segmask = mask_range(s_lane, e_lane);
            const bool head = act && lane == s_lane;
            const uint64_t p = (uint64_t)gstart + (lane - s_lane);
            const uint64_t key = act ? rec_u64(keys + p) : 0ull;
            const uint32_t ord = act ? rec_u32(vals + p) : 0u;
            const ProtMeta m = act ? load_meta(meta, ord) : make_uint2(0, 0);           // len, func
            const uint32_t f = m.y;
            const uint32_t off = sigk_key_offset(key);

            // function vote (func_count + arg-max, tcc:203, :228-248).  A group with two
            // different functions and fewer than 5 records cannot reach 80 % (1/2, 2/3, 3/4),
            // so the bit-sliced vote only runs when a mixed group has 5 or more records.
            uint32_t cand = __shfl_sync(FULL, f, s_lane);
            const bool mixed = (__ballot_sync(FULL, act && f != cand) & segmask) != 0;
            if (__any_sync(FULL, act && mixed && cnt >= 5)) {
                uint32_t c = 0;
#pragma unroll
                for (int b = 0; b < 16; ++b) {
                    const unsigned bal = __ballot_sync(FULL, act && ((f >> b) & 1u)) & segmask;
                    c |= (2u * (uint32_t)__popc(bal) > cnt ? 1u : 0u) << b;
                }
                if (mixed) cand = c;
            }
            const bool is_best = act && f == cand;
            const unsigned best_mask = __ballot_sync(FULL, is_best) & segmask;
            const uint32_t best_count = __popc(best_mask);
            const bool keep = act && !(mixed && cnt < 5) && keep_rule(best_count, cnt);

            // avg_from_end: rank cnt/2 of the group's offsets (tcc:273, :281-282), radix select
            // over the bits on which some group of the window disagrees
            const uint32_t off_head = __shfl_sync(FULL, off, s_lane);
            uint32_t vary = __reduce_or_sync(FULL, act ? (off ^ off_head) : 0u);
            uint32_t sel = off & ~vary;
            {
                unsigned cm = segmask;
                uint32_t rk = cnt >> 1;
                while (vary) {
                    const int b = 31 - __clz(vary);
                    vary &= ~(1u << b);
                    const unsigned onesb = __ballot_sync(FULL, act && ((off >> b) & 1u)) & cm;
                    const unsigned zeros = cm & ~onesb;
                    const uint32_t cz = __popc(zeros);
                    if (rk < cz) cm = zeros;
                    else { rk -= cz; cm = onesb; sel |= 1u << b; }
                }
            }
            // sum of the best function's lengths mod 65536 (sum_impl<unsigned short>)
            uint32_t x = is_best ? (m.x & 0xFFFFu) : 0u;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(FULL, x, o);
                if (lane >= s_lane + (unsigned)o) x += y;
            }
            const uint32_t S = __shfl_sync(FULL, x, (e_lane - 1u) & 31u) & 0xFFFFu;

            // two best items: median stays 0 (heights[2] untouched) and the variance recurrence collapses to
            // var = (x2 - S/2)^2 with x2 the second sample visited = the earlier of the two in sorted order;
            // every operation is exact in double, so no division is needed.
            uint32_t var2 = 0;
            if (order_stats && __any_sync(FULL, keep && head && best_count == 2)) {
                const uint32_t x2 = __shfl_sync(FULL, m.x, (__ffs(best_mask) - 1) & 31);
                const double tmp = __dsub_rn((double)x2, __dmul_rn((double)S, 0.5));
                var2 = best_count == 2 ? u16_from_double(__dmul_rn(tmp, tmp)) : 0u;
            }
            if (act && !keep) count_rejected(prot_rejected, ord);

            // Three or more best items: if they all have the same length L and best_count * L < 65536
            // (the 16-bit sum has not wrapped), every P^2 height is L and every variance step adds
            // (L - (n L)/n)^2 = 0 exactly, so median = L and var = 0 without walking the group.
            const uint32_t len0 = __shfl_sync(FULL, m.x, (__ffs(best_mask) - 1) & 31);
            const bool uneven = (__ballot_sync(FULL, is_best && m.x != len0) & segmask) != 0;
            const bool closed = !uneven && (uint64_t)best_count * len0 < 65536ull;
            if (best_count >= 3 && closed) var2 = 0;
            const uint32_t median_now = (best_count >= 3 && closed) ? len0 : 0u;
            const bool walk = keep && head && order_stats && best_count >= 3 && !closed;
            const unsigned wb = __ballot_sync(FULL, walk);
            if (wb) work_reserve(wc, (uint32_t)__popc(wb), work, n_work);
            if (head) {
                if (keep) {
                    const uint64_t code = sigk_key_code(key);
                    // u16((double)S / n) = floor(S / best_count), S < 65536, best_count <= 32: float quotient, fixed up
                    uint32_t mean = (uint32_t)__float2uint_rz(__fmul_rn((float)S, __frcp_rn((float)best_count)));
                    if (mean * best_count > S) --mean;
                    else if ((mean + 1u) * best_count <= S) ++mean;
                    rows[grow] = make_uint4((uint32_t)code, (uint32_t)(code >> 32) | (sel << 11), cand | (mean << 16),
                                            (order_stats ? median_now : 0u) | (var2 << 16));
                    if (walk) work[wc.base + __popc(wb & mask_lt(lane))] = OrderWork{grow, (uint32_t)p, cnt};
                } else {
                    rows[grow] = make_uint4(0u, 0u, 0xFFFFu, 0u);
                    atomicAdd(rej_tile + (grow >> SQ_TILE_SHIFT), 1u);          // the squeeze's per-tile tombstone count
                }
            }
            if (wb) { const uint32_t k = __popc(wb); wc.base += k; wc.free -= k; }
            base += wlen;
            j0 += nd;
        }
      }
    }
    work_flush(wc, work);
}

// ---- groups of more than REDUCE_GIANT records: reduce_long_segment with the whole CTA -----------------------------
// On the Zipf set the largest family leaves ~300 groups of 250 K records; one warp walking such a group nine times
// (count, then one walk per varying offset bit) was the whole tail of group_reduce_kernel (19 ms at config 4).
constexpr int GT_THREADS = 512;
constexpr int GT_WARPS = GT_THREADS / 32;
template <typename T, typename Op>
SIGK_D T giant_block_reduce(T v, T *s_red, Op op) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(FULL, v, o));
    __syncthreads();                    // (s_red may still be read from the previous reduction)
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    T r = s_red[0];
    for (int w = 1; w < GT_WARPS; ++w) r = op(r, s_red[w]);
    return r;
}
template <typename MetaT>
__global__ void __launch_bounds__(GT_THREADS)
giant_reduce_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, const MetaT *__restrict__ meta,
                    const OrderWork *__restrict__ giant, const uint32_t *__restrict__ n_giant, uint32_t *__restrict__ next_giant,
                    uint4 *__restrict__ rows, OrderWork *__restrict__ work, uint32_t *__restrict__ n_work,
                    OrderWork *__restrict__ work_long, uint32_t *__restrict__ n_work_long, uint32_t *__restrict__ prot_rejected,
                    uint32_t *__restrict__ rej_tile, int order_stats) {
    __shared__ uint32_t s_red[GT_WARPS];
    __shared__ uint32_t s_ones[16];
    __shared__ uint32_t s_next;
    const uint32_t total = *n_giant;
    auto add = [](uint32_t a, uint32_t b) { return a + b; };
    auto bor = [](uint32_t a, uint32_t b) { return a | b; };
    auto mn = [](uint32_t a, uint32_t b) { return min(a, b); };
    auto mx = [](uint32_t a, uint32_t b) { return max(a, b); };
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_next = atomicAdd(next_giant, 1u);
        if (threadIdx.x < 16) s_ones[threadIdx.x] = 0;
        __syncthreads();
        const uint32_t i = s_next;
        if (i >= total) return;
        const OrderWork d = giant[i];
        const uint64_t start = d.start;
        const uint32_t n = d.count;
        // the same three walks as reduce_long_segment, strided over the CTA
        const uint32_t off0 = sigk_key_offset(keys[start]);
        uint32_t cand = load_meta(meta, vals[start]).y;
        uint32_t best = 0, S = 0, vary = 0, lmin = 0xFFFFFFFFu, lmax = 0, mixed = 0;
        for (uint32_t j = threadIdx.x; j < n; j += GT_THREADS) {
            const ProtMeta m = load_meta(meta, vals[start + j]);
            if (m.y == cand) { ++best; S += m.x; lmin = min(lmin, m.x); lmax = max(lmax, m.x); }
            else mixed = 1;
            vary |= sigk_key_offset(keys[start + j]) ^ off0;
        }
        mixed = giant_block_reduce(mixed, s_red, bor);
        if (mixed) {
            uint32_t ones[16];
#pragma unroll
            for (int b = 0; b < 16; ++b) ones[b] = 0;
            for (uint32_t j = threadIdx.x; j < n; j += GT_THREADS) {
                const uint32_t f = load_meta(meta, vals[start + j]).y;
#pragma unroll
                for (int b = 0; b < 16; ++b) ones[b] += (f >> b) & 1u;
            }
#pragma unroll
            for (int b = 0; b < 16; ++b) {
                uint32_t v = ones[b];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
                if ((threadIdx.x & 31u) == 0) atomicAdd(&s_ones[b], v);
            }
            __syncthreads();
            uint32_t voted = 0;
#pragma unroll
            for (int b = 0; b < 16; ++b) voted |= (2ull * s_ones[b] > n ? 1u : 0u) << b;
            if (voted != cand) {
                cand = voted;
                best = 0; S = 0; lmin = 0xFFFFFFFFu; lmax = 0;
                for (uint32_t j = threadIdx.x; j < n; j += GT_THREADS) {
                    const ProtMeta m = load_meta(meta, vals[start + j]);
                    if (m.y == cand) { ++best; S += m.x; lmin = min(lmin, m.x); lmax = max(lmax, m.x); }
                }
            }
        }
        best = giant_block_reduce(best, s_red, add);
        S = giant_block_reduce(S, s_red, add);
        vary = giant_block_reduce(vary, s_red, bor);
        lmin = giant_block_reduce(lmin, s_red, mn);
        lmax = giant_block_reduce(lmax, s_red, mx);
        const bool keep = keep_rule(best, n);
        if (!keep) {
            for (uint32_t j = threadIdx.x; j < n; j += GT_THREADS) count_rejected(prot_rejected, vals[start + j]);
            if (threadIdx.x == 0) {
                rows[d.row] = make_uint4(0u, 0u, 0xFFFFu, 0u);
                atomicAdd(rej_tile + (d.row >> SQ_TILE_SHIFT), 1u);
            }
            continue;
        }
        const uint32_t mean = (S & 0xFFFFu) / best;
        const bool closed = lmin == lmax && (uint64_t)best * lmin < 65536ull;
        // offset of rank n/2 by radix select over the varying bits, most significant first
        uint32_t rank = n / 2, prefix = off0 & ~vary, pmask = ~vary & 0xFFFFu;
        while (vary) {
            const int b = 31 - __clz(vary);
            uint32_t zeros = 0;
            for (uint32_t j = threadIdx.x; j < n; j += GT_THREADS) {
                const uint32_t off = sigk_key_offset(keys[start + j]);
                zeros += ((off & pmask) == prefix && !((off >> b) & 1u)) ? 1u : 0u;
            }
            zeros = giant_block_reduce(zeros, s_red, add);
            if (rank >= zeros) { rank -= zeros; prefix |= 1u << b; }
            pmask |= 1u << b;
            vary &= ~(1u << b);
        }
        if (threadIdx.x == 0) {
            const uint64_t code = sigk_key_code(keys[start]);
            rows[d.row] = make_uint4((uint32_t)code, (uint32_t)(code >> 32) | (prefix << 11), cand | (mean << 16), (order_stats && closed) ? lmin : 0u);
            if (order_stats && !closed) {           // best_count >= 0.8 * 4096 here: the ordered walk is needed
                if (n > ORD_LONG) work_long[atomicAdd(n_work_long, 1u)] = d;
                else work[atomicAdd(n_work, 1u)] = d;
            }
        }
    }
}

// ---- squeeze: drop the tombstones, keep k-mer order, expand rows into the table columns
// Only rejected groups of two or more records leave tombstones, and group_reduce_kernel counts them per
// squeeze tile as it writes them; one small scan later every tile knows where its kept rows go, so the
// tiles are independent of each other (no chained scan, no tickets).

// two residues at a time: pair_ascii[20 a + b] = letter(a) | letter(b) << 8 (upper case; the case mask is ORed in
// afterwards).  A compile-time table: nothing to initialise, nothing to order against the build streams.
struct PairAscii { uint16_t v[400]; };
constexpr PairAscii make_pair_ascii() {
    PairAscii t{};
    const char aa[21] = "ACDEFGHIKLMNPQRSTVWY";
    for (int i = 0; i < 400; ++i) t.v[i] = (uint16_t)((unsigned char)aa[i / 20] | ((unsigned char)aa[i % 20] << 8));
    return t;
}
__device__ const PairAscii g_pair_ascii = make_pair_ascii();
// group code (code35 << 8 | mask8) -> the k-mer's 8 ASCII bytes, first residue in the low byte
SIGK_D uint64_t code_to_ascii_pairs(uint64_t gcode) {
    const uint64_t code = gcode >> SIGK_MASK_BITS;
    const uint32_t mask = (uint32_t)gcode & 0xFFu;
    const uint32_t hi = (uint32_t)(code / 160000ull);                       // 20^4
    const uint32_t lo = (uint32_t)(code - (uint64_t)hi * 160000ull);
    uint32_t w0 = __ldg(g_pair_ascii.v + hi / 400u) | ((uint32_t)__ldg(g_pair_ascii.v + hi % 400u) << 16);
    uint32_t w1 = __ldg(g_pair_ascii.v + lo / 400u) | ((uint32_t)__ldg(g_pair_ascii.v + lo % 400u) << 16);
    if (mask) {                                                              // rare: 0x20 into the lower-case positions
        const uint64_t m = sigk_spread_mask(mask);
        w0 |= (uint32_t)m; w1 |= (uint32_t)(m >> 32);
    }
    return (uint64_t)w0 | ((uint64_t)w1 << 32);
}

// rej_before[t] = tombstones in the squeeze tiles before t; n_kept = groups - all tombstones.  One block.
__global__ void __launch_bounds__(SCAN_THREADS)
tombstone_scan_kernel(const uint64_t *__restrict__ n_seg_ptr, const uint32_t *__restrict__ rej_tile,
                      uint32_t *__restrict__ rej_before, uint64_t *__restrict__ n_kept_out) {
    __shared__ uint32_t s_part[SCAN_THREADS / 32 + 1];
    const uint64_t n_seg = *n_seg_ptr;
    const uint32_t total = block_scan_tiles<uint32_t>(rej_tile, rej_before, (n_seg + SQ_TILE - 1) / SQ_TILE, s_part);
    if (threadIdx.x == 0) *n_kept_out = n_seg - total;
}

#ifndef SIGK_SQ_MIN_BLOCKS
#define SIGK_SQ_MIN_BLOCKS 6
#endif
__global__ void __launch_bounds__(SQ_THREADS, SIGK_SQ_MIN_BLOCKS)
squeeze_rows_kernel(const uint4 *__restrict__ rows, const uint64_t *__restrict__ n_seg_ptr, KeptColumns out,
                    const uint32_t *__restrict__ rej_before, uint64_t *__restrict__ n_side_kept) {
    __shared__ uint32_t s_scan[SQ_WARPS + 2];
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint64_t n_seg = *n_seg_ptr;
    const uint32_t tile = blockIdx.x;
    const uint64_t tile_start = (uint64_t)tile * SQ_TILE;
    if (tile_start >= n_seg) return;
    const uint64_t base = tile_start - __ldg(rej_before + tile);

    uint4 row[SQ_ITEMS];
    unsigned ball[SQ_ITEMS];
    uint32_t warp_total = 0;
    const uint64_t wbase = tile_start + (uint64_t)warp * (SQ_ITEMS * 32);
#pragma unroll
    for (int i = 0; i < SQ_ITEMS; ++i) {
        const uint64_t s = wbase + i * 32 + lane;
        row[i] = s < n_seg ? __ldcs(rows + s) : make_uint4(0u, 0u, 0xFFFFu, 0u);
        ball[i] = __ballot_sync(FULL, (row[i].z & 0xFFFFu) != 0xFFFFu);
        warp_total += __popc(ball[i]);
    }
    uint32_t total;
    const uint32_t excl = block_exclusive_scan<SQ_THREADS>(lane == 0 ? warp_total : 0u, s_scan, &total);
    uint32_t run = __shfl_sync(FULL, excl, 0);
    uint32_t side = 0;                          // kept rows of the side run (case mask != 0): the table's second section
#pragma unroll
    for (int i = 0; i < SQ_ITEMS; ++i) {
        if ((ball[i] >> lane) & 1u) {
            side += (row[i].x & 0xFFu) ? 1u : 0u;
            const uint64_t o = base + run + __popc(ball[i] & mask_lt(lane));
            out.kmer[o] = code_to_ascii_pairs((uint64_t)row[i].x | ((uint64_t)(row[i].y & 0x7FFu) << 32));
            out.avg_from_end[o] = (uint16_t)(row[i].y >> 11);
            out.function_index[o] = (uint16_t)(row[i].z & 0xFFFFu);
            out.mean[o] = (uint16_t)(row[i].z >> 16);
            out.median[o] = (uint16_t)(row[i].w & 0xFFFFu);
            out.var[o] = (uint16_t)(row[i].w >> 16);
        }
        run += __popc(ball[i]);
    }
    if (__any_sync(FULL, side != 0)) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) side += __shfl_xor_sync(FULL, side, o);
        if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long *>(n_side_kept), (unsigned long long)side);
    }
}

// ---- order-dependent columns: median (P^2) and var (iterative) of the kept groups in `work`.
// The recurrences are sequential per group, so a lane owns one group at a time and a lane
// that finishes takes the next entry of its warp's block at once (groups differ in length by
// orders of magnitude; waiting for the slowest lane left 13 % of the lanes busy in the v2 profile).
constexpr int ORD_THREADS = 128;
constexpr int ORD_BLOCK = 64;       // work entries a warp takes per fetch

// Groups of more than ORD_LONG records: a warp per group and per recurrence.  The recurrences stay
// sequential, but the warp fetches 32 records at a time (coalesced values, 32 meta gathers in flight,
// the next batch prefetched) and every lane runs the same accumulator on shuffled samples, so a group
// does not pay two dependent memory latencies per record.  What is left is the latency of the
// dependent double-precision chain, sample after sample (config 4: 250 K samples in the largest
// groups, one chain of ~1200 cycles per sample in the round-1 kernel), so the median (P^2 markers)
// and the variance — independent recurrences — go to two different warps (kind 0 and 1 of the
// group's work entry), and the variance warp computes everything that does not depend on the running
// variance for 32 samples at once: which samples count, their running count n_i and wrapped sum S_i
// (prefix scans), the term tmp_i^2/(n_i-1) and the refined reciprocal of n_i.
template <typename MetaT>
SIGK_D void order_stats_long_group(const uint32_t *__restrict__ vals, const MetaT *__restrict__ meta, const OrderWork w, const uint32_t kind,
                                   uint4 *__restrict__ rows) {
    const unsigned lane = threadIdx.x & 31u;
    __shared__ double s_stage[2][ORD_THREADS / 32][32];
    double *st_a = s_stage[0][threadIdx.x >> 5], *st_b = s_stage[1][threadIdx.x >> 5];
    const uint32_t cand = rows[w.row].z & 0xFFFFu;
    LengthAcc acc;          // every lane carries the same state
    // newest first: batch b covers records count-1-32b-lane
    int64_t j = (int64_t)w.count - 1 - (int64_t)lane;
    ProtMeta m = j >= 0 ? load_meta(meta, __ldg(vals + (uint64_t)w.start + j)) : make_uint2(0, 0xFFFFFFFFu);
    for (int64_t left = w.count; left > 0; left -= 32) {
        const int64_t jn = j - 32;
        const ProtMeta mn = (left > 32 && jn >= 0) ? load_meta(meta, __ldg(vals + (uint64_t)w.start + jn)) : make_uint2(0, 0xFFFFFFFFu);
        const bool match = m.y == cand;
        const unsigned mb = __ballot_sync(FULL, match);
        if (mb) {
            // the samples that count, compacted into the warp's staging: the chain below reads them in order,
            // one shared-memory load per sample (issued a sample ahead) instead of a find-bit and a shuffle
            const uint32_t cnt = __popc(mb), slot = __popc(mb & mask_lt(lane));
            if (kind == 0) {
                if (match) st_a[slot] = (double)m.x;
                __syncwarp();
                double xn = st_a[0];
                for (uint32_t k = 0; k < cnt; ++k) {
                    const double x = xn;
                    xn = st_a[(k + 1) & 31u];
                    acc.n += 1;
                    acc.quantile_step(x);                                // acc(item.protein_length), tcc:271
                }
            } else {
                uint32_t sx = match ? m.x : 0u;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(FULL, sx, o);
                    if (lane >= (unsigned)o) sx += y;
                }
                const uint32_t n_i = acc.n + slot + 1;
                const uint32_t S_i = (acc.S + sx) & 0xFFFFu;
                if (match) {
                    st_a[slot] = n_i > 1 ? LengthAcc::variance_term(m.x, S_i, n_i) : 0.0;
                    st_b[slot] = recip_refined((double)n_i);
                }
                __syncwarp();
                double tn = st_a[0], yn = st_b[0];
                for (uint32_t k = 0; k < cnt; ++k) {
                    const double t = tn, y = yn;
                    tn = st_a[(k + 1) & 31u]; yn = st_b[(k + 1) & 31u];
                    acc.n += 1;
                    if (acc.n > 1) acc.variance_step_pre((double)acc.n, (double)(acc.n - 1), y, t);
                }
                acc.S = (acc.S + __shfl_sync(FULL, sx, 31)) & 0xFFFFu;
            }
            __syncwarp();
        }
        m = mn;
        j = jn;
    }
    if (lane == 0) {
        uint16_t *half = reinterpret_cast<uint16_t *>(&rows[w.row].w);   // median: low half, var: high half (tcc:278-279)
        if (kind == 0) half[0] = (uint16_t)u16_from_double(acc.q2);
        else half[1] = (uint16_t)u16_from_double(acc.var);
    }
}

// One kernel for both lists: every warp first serves the long list (two entries per group), then
// joins the per-lane walk of the ordinary one, so the tail of the long groups runs beside it.
template <typename MetaT>
__global__ void __launch_bounds__(ORD_THREADS)
order_stats_kernel(const uint32_t *__restrict__ vals, const MetaT *__restrict__ meta, const OrderWork *__restrict__ work,
                   const uint32_t *__restrict__ n_work, uint32_t *__restrict__ next, const OrderWork *__restrict__ work_long,
                   const uint32_t *__restrict__ n_work_long, uint32_t *__restrict__ next_long, uint4 *__restrict__ rows) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t total_long = *n_work_long * 2u;
    for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = atomicAdd(next_long, 1u);
        i = __shfl_sync(FULL, i, 0);
        if (i >= total_long) break;
        order_stats_long_group(vals, meta, work_long[i >> 1], i & 1u, rows);
    }
    const uint32_t total = *n_work;
    // Blocks of ORD_BLOCK entries are handed out dynamically, and a warp takes its next block as soon as
    // one of its lanes is idle — not when all of them are: entries of one family sit together in the list
    // and range from 3 to tens of thousands of samples, so waiting for a block's longest group idled
    // three lanes in four on the Zipf set.
    uint32_t cursor = 0, b1 = 0;
    bool drained = false;                       // the global list is exhausted
    bool active = false;
    uint32_t row = 0, cand = 0, left = 0;
    uint64_t start = 0;
    ProtMeta m_next = make_uint2(0, 0);         // meta of the sample to visit next, already in flight
    LengthAcc acc;
    for (;;) {
        const unsigned idle = __ballot_sync(FULL, !active);
        if (idle) {
            if (cursor >= b1 && !drained) {
                uint32_t nb = 0;
                if (lane == 0) nb = atomicAdd(next, (uint32_t)ORD_BLOCK);
                nb = __shfl_sync(FULL, nb, 0);
                if (nb >= total) drained = true;
                else { cursor = nb; b1 = (nb + ORD_BLOCK < total) ? nb + ORD_BLOCK : total; }
            }
            if (cursor < b1) {
                const uint32_t take = min((uint32_t)__popc(idle), b1 - cursor);
                const uint32_t r = __popc(idle & mask_lt(lane));
                if (!active && r < take) {
                    const OrderWork w = work[cursor + r];
                    if (w.count) {                                         // count 0: an unused reserved slot
                        row = w.row; start = w.start; left = w.count;
                        cand = rows[w.row].z & 0xFFFFu;
                        m_next = load_meta(meta, __ldg(vals + start + left - 1));
                        acc = LengthAcc();
                        active = true;
                    }
                }
                cursor += take;
            }
        }
        if (!__any_sync(FULL, active)) { if (drained && cursor >= b1) break; else continue; }
        if (active) {
            // newest first: the multimap iterates a key's items in reverse insertion order
            --left;
            const ProtMeta m = m_next;
            if (left) m_next = load_meta(meta, __ldg(vals + start + left - 1));
            if (m.y == cand) acc.push(m.x);                            // acc(item.protein_length), tcc:271
            if (left == 0) {
                rows[row].w = u16_from_double(acc.q2) | (u16_from_double(acc.var) << 16);   // tcc:278-279
                active = false;
            }
        }
    }
}

// ---- distinct_functions[best]++ per kept k-mer (tcc:286) --------------------------------------
// One global atomic per kept row (hundreds of millions on a few ten thousand counters) is what bounded the
// run-length kernel, so the tally is taken afterwards from the function_index column of the finished table:
// every block owns a range of 32768 counters in shared memory, walks its share of the column with 16-byte
// loads and flushes what it counted once.
constexpr int FH_THREADS = 1024;
constexpr int FH_BINS = 32768;

__global__ void __launch_bounds__(FH_THREADS)
function_histogram_kernel(const uint16_t *__restrict__ func, const uint64_t *__restrict__ n_ptr, int n_ranges,
                          uint32_t *__restrict__ distinct_functions) {
    extern __shared__ uint32_t fh_bins[];
    const uint32_t range = blockIdx.x % (uint32_t)n_ranges, chunk = blockIdx.x / (uint32_t)n_ranges;
    const uint32_t n_chunks = gridDim.x / (uint32_t)n_ranges;
    if (chunk >= n_chunks) return;
    for (int b = threadIdx.x; b < FH_BINS; b += FH_THREADS) fh_bins[b] = 0;
    __syncthreads();
    const uint64_t n = *n_ptr;
    auto tally = [&](uint32_t f) { if ((f >> 15) == range) atomicAdd(&fh_bins[f & (FH_BINS - 1)], 1u); };
    // scalar head up to 16-byte alignment, 8 rows per load in the body, scalar tail
    const uint64_t mis = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(func) & 15u)) & 15u;
    const uint64_t head = min(n, mis / 2);
    const uint64_t n8 = (n - head) / 8;
    const uint4 *body = reinterpret_cast<const uint4 *>(func + head);
    for (uint64_t i = (uint64_t)chunk * FH_THREADS + threadIdx.x; i < n8; i += (uint64_t)n_chunks * FH_THREADS) {
        const uint4 v = __ldg(body + i);
        tally(v.x & 0xFFFFu); tally(v.x >> 16); tally(v.y & 0xFFFFu); tally(v.y >> 16);
        tally(v.z & 0xFFFFu); tally(v.z >> 16); tally(v.w & 0xFFFFu); tally(v.w >> 16);
    }
    if (chunk == 0) {
        for (uint64_t i = threadIdx.x; i < head; i += FH_THREADS) tally(func[i]);
        for (uint64_t i = head + n8 * 8 + threadIdx.x; i < n; i += FH_THREADS) tally(func[i]);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < FH_BINS; b += FH_THREADS)
        if (fh_bins[b]) atomicAdd(distinct_functions + (size_t)range * FH_BINS + b, fh_bins[b]);
}

// seq_bitmap bit seq_id[i] = protein i has an occurrence in a kept group
__global__ void signature_flags_kernel(const uint32_t *__restrict__ prot_windows, const uint32_t *__restrict__ prot_rejected,
                                       const uint32_t *__restrict__ seq_id, uint32_t n_prot, uint32_t *__restrict__ bitmap) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_prot) return;
    if (prot_windows[i] > prot_rejected[(size_t)i << c_rej_shift]) {
        const uint32_t sid = seq_id[i];
        atomicOr(bitmap + (sid >> 5), 1u << (sid & 31u));
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

template <typename MetaT>
__global__ void protein_meta_kernel(const uint64_t *__restrict__ starts, const uint16_t *__restrict__ func,
                                    const uint32_t *__restrict__ seq_id, uint32_t n_prot, MetaT *__restrict__ meta,
                                    uint32_t *__restrict__ seqs_with_func) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_prot) return;
    (void)seq_id;
    const uint32_t len = (uint32_t)(starts[i + 1] - starts[i]);
    if (sizeof(MetaT) == sizeof(uint32_t)) reinterpret_cast<uint32_t *>(meta)[(size_t)i << c_meta_shift] = len | ((uint32_t)func[i] << 16);     // len < 65 535 (caller)
    else reinterpret_cast<ProtMeta *>(meta)[(size_t)i << c_meta_shift] = make_uint2(len, func[i]);
    if (seqs_with_func) atomicAdd(seqs_with_func + func[i], 1u);          // seqs_with_func[function_index]++, tcc:160
}

}  // namespace

static int reduce_grid(int sm_count) { return sm_count * 12; }
size_t reduce_group_entries(uint64_t capacity, int) {
    return (size_t)(capacity / 2 + 2);      // a listed group has >= 2 records
}
size_t reduce_long_group_entries(uint64_t capacity) { return (size_t)(capacity / 33 + 2); }
size_t reduce_giant_entries(uint64_t capacity) { return (size_t)(capacity / REDUCE_GIANT + 2); }
size_t reduce_work_entries(uint64_t capacity, int sm_count) {
    // A walked group has >= 3 records.  Slots are reserved in blocks of WORK_BLOCK and a block's unused tail is
    // dropped when the next window needs more than it has left (a window lists at most 10 groups), so a block
    // holds at least WORK_BLOCK - 9 live entries: size the list for that padding, plus one open block per warp.
    const size_t groups = (size_t)(capacity / 3 + 2);
    // (+ one slot per giant group: giant_reduce_kernel appends singly)
    return groups + groups * 9 / (WORK_BLOCK - 9) + WORK_BLOCK + 2 * (size_t)reduce_grid(sm_count) * (RED_THREADS / 32) * WORK_BLOCK + reduce_giant_entries(capacity);
}
size_t reduce_long_work_entries(uint64_t capacity) { return (size_t)(capacity / ORD_LONG + 2); }

cudaError_t reduce_configure(int meta_shift, int rej_shift) {
    const uint32_t s = (uint32_t)meta_shift, r = (uint32_t)rej_shift;
    cudaError_t e = cudaMemcpyToSymbol(c_rej_shift, &r, sizeof r);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_meta_shift, &s, sizeof s);
}

cudaError_t launch_protein_meta(const uint64_t *starts, const uint16_t *func, const uint32_t *seq_id, uint32_t n_prot,
                                MetaTable meta, uint64_t first, uint32_t *seqs_with_func, cudaStream_t stream) {
    if (n_prot == 0) return cudaSuccess;
    const unsigned grid = (n_prot + 255) / 256;
    if (meta.compact)
        protein_meta_kernel<uint32_t><<<grid, 256, 0, stream>>>(starts, func, seq_id, n_prot, static_cast<uint32_t *>(meta.p) + (first << meta.shift), seqs_with_func);
    else
        protein_meta_kernel<ProtMeta><<<grid, 256, 0, stream>>>(starts, func, seq_id, n_prot, static_cast<ProtMeta *>(meta.p) + (first << meta.shift), seqs_with_func);
    return cudaGetLastError();
}

ReduceScratch reduce_scratch(uint64_t *words, uint64_t capacity) {
    const uint64_t b = reduce_batches(capacity) + 1, q = squeeze_tiles(capacity) + 1;
    ReduceScratch sx;
    sx.tile_counts = words;
    sx.tile_base = words + b;
    sx.tile_open = reinterpret_cast<uint32_t *>(words + 2 * b);
    sx.tile_first = sx.tile_open + b;
    sx.rej_tile = reinterpret_cast<uint32_t *>(words + 3 * b);
    sx.rej_before = sx.rej_tile + q;
    return sx;
}

template <typename MetaT>
static cudaError_t segment_reduce_impl(const uint64_t *keys, const uint32_t *vals, const uint64_t *n_ptr, uint64_t capacity,
                                       const MetaT *meta, uint4 *rows, const ReduceLists &l, uint32_t *prot_rejected,
                                       const ReduceScratch &sx, uint64_t *n_seg_out, int order_stats, int sm_count, cudaStream_t stream,
                                       cudaEvent_t after_count, cudaEvent_t after_emit) {
    const uint64_t grid = reduce_grid(sm_count);
    const uint64_t tiles = reduce_batches(capacity);
    head_tile_kernel<false, MetaT><<<(unsigned)tiles, HS_THREADS, 0, stream>>>(keys, vals, n_ptr, meta, rows, l.groups, l.long_groups, l.n_long, sx);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    tile_scan_kernel<<<1, SCAN_THREADS, 0, stream>>>(n_ptr, sx, n_seg_out, l.n_groups);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (after_count) cudaEventRecord(after_count, stream);
    head_tile_kernel<true, MetaT><<<(unsigned)tiles, HS_THREADS, 0, stream>>>(keys, vals, n_ptr, meta, rows, l.groups, l.long_groups, l.n_long, sx);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    resolve_open_kernel<MetaT><<<(unsigned)((tiles + 255) / 256), 256, 0, stream>>>(keys, vals, n_ptr, meta, sx, rows, l.groups, l.n_groups,
                                                                                     l.long_groups, l.n_long);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (after_emit) cudaEventRecord(after_emit, stream);
    group_reduce_kernel<MetaT><<<(unsigned)grid, RED_THREADS, 0, stream>>>(keys, vals, meta, l.groups, l.n_groups, l.next_group, l.long_groups,
                                                                           l.n_long, l.next_long, rows, l.work, l.n_work, l.work_long,
                                                                           l.n_work_long, l.giant, l.n_giant, prot_rejected, sx.rej_tile, order_stats);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    giant_reduce_kernel<MetaT><<<(unsigned)sm_count * 2, GT_THREADS, 0, stream>>>(keys, vals, meta, l.giant, l.n_giant, l.next_giant, rows, l.work, l.n_work,
                                                                                  l.work_long, l.n_work_long, prot_rejected, sx.rej_tile, order_stats);
    return cudaGetLastError();
}

cudaError_t launch_segment_reduce(const uint64_t *keys, const uint32_t *vals, const uint64_t *n_ptr, uint64_t capacity,
                                  MetaTable meta, uint4 *rows, const ReduceLists &l, uint32_t *prot_rejected,
                                  uint64_t *scratch_words, uint64_t *n_seg_out, int order_stats, int sm_count, cudaStream_t stream,
                                  cudaEvent_t after_count, cudaEvent_t after_emit) {
    if (capacity == 0) return cudaSuccess;
    const ReduceScratch sx = reduce_scratch(scratch_words, capacity);
    return meta.compact ? segment_reduce_impl(keys, vals, n_ptr, capacity, static_cast<const uint32_t *>(meta.p), rows, l, prot_rejected, sx,
                                              n_seg_out, order_stats, sm_count, stream, after_count, after_emit)
                        : segment_reduce_impl(keys, vals, n_ptr, capacity, static_cast<const ProtMeta *>(meta.p), rows, l, prot_rejected, sx,
                                              n_seg_out, order_stats, sm_count, stream, after_count, after_emit);
}

template <typename MetaT>
static cudaError_t order_stats_impl(const uint32_t *vals, const MetaT *meta, const OrderWork *work, const uint32_t *n_work,
                                    uint32_t *next_work, const OrderWork *work_long, const uint32_t *n_work_long, uint32_t *next_long,
                                    uint4 *rows, int sm_count, cudaStream_t stream) {
    order_stats_kernel<MetaT><<<sm_count * 16, ORD_THREADS, 0, stream>>>(vals, meta, work, n_work, next_work, work_long, n_work_long, next_long, rows);
    return cudaGetLastError();
}

// the inlined division of length_acc.cuh beside the toolkit's, for tests/test_gpu_parity.py
__global__ void ddiv_check_kernel(const double *__restrict__ a, const double *__restrict__ b, uint64_t n, double *__restrict__ inl,
                                  double *__restrict__ lib) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    inl[i] = ddiv_inline(a[i], b[i]);
    lib[i] = __ddiv_rn(a[i], b[i]);
}
cudaError_t launch_ddiv_check(const double *a, const double *b, uint64_t n, double *inl, double *lib, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    ddiv_check_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(a, b, n, inl, lib);
    return cudaGetLastError();
}

cudaError_t launch_order_stats(const uint32_t *vals, MetaTable meta, const OrderWork *work, const uint32_t *n_work,
                               uint32_t *next_work, const OrderWork *work_long, const uint32_t *n_work_long, uint32_t *next_long,
                               uint64_t capacity, uint4 *rows, int sm_count, cudaStream_t stream) {
    if (capacity == 0) return cudaSuccess;
    return meta.compact ? order_stats_impl(vals, static_cast<const uint32_t *>(meta.p), work, n_work, next_work, work_long, n_work_long,
                                           next_long, rows, sm_count, stream)
                        : order_stats_impl(vals, static_cast<const ProtMeta *>(meta.p), work, n_work, next_work, work_long, n_work_long,
                                           next_long, rows, sm_count, stream);
}

cudaError_t launch_squeeze_rows(const uint4 *rows, const uint64_t *n_seg_ptr, uint64_t capacity, KeptColumns out,
                                uint64_t *scratch_words, uint64_t *n_kept_out, uint64_t *n_side_kept, cudaStream_t stream) {
    if (capacity == 0) return cudaSuccess;
    const ReduceScratch sx = reduce_scratch(scratch_words, capacity);
    tombstone_scan_kernel<<<1, SCAN_THREADS, 0, stream>>>(n_seg_ptr, sx.rej_tile, sx.rej_before, n_kept_out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    squeeze_rows_kernel<<<(unsigned)squeeze_tiles(capacity), SQ_THREADS, 0, stream>>>(rows, n_seg_ptr, out, sx.rej_before, n_side_kept);
    return cudaGetLastError();
}

cudaError_t launch_function_histogram(const uint16_t *function_index, const uint64_t *n_kept_ptr, uint64_t capacity,
                                      uint32_t max_function, uint32_t *distinct_functions, int sm_count, cudaStream_t stream) {
    if (capacity == 0) return cudaSuccess;
    // per device, so set on every launch (a process may drive several devices through several handles)
    cudaError_t attr = cudaFuncSetAttribute(function_histogram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FH_BINS * (int)sizeof(uint32_t));
    if (attr != cudaSuccess) return attr;
    const int n_ranges = max_function >= (uint32_t)FH_BINS ? 2 : 1;
    const uint64_t want = (capacity / 8 + FH_THREADS - 1) / FH_THREADS + 1;
    const unsigned chunks = (unsigned)std::min<uint64_t>((uint64_t)sm_count, want);
    function_histogram_kernel<<<chunks * n_ranges, FH_THREADS, FH_BINS * sizeof(uint32_t), stream>>>(function_index, n_kept_ptr, n_ranges,
                                                                                                  distinct_functions);
    return cudaGetLastError();
}

cudaError_t launch_signature_flags(const uint32_t *prot_windows, const uint32_t *prot_rejected, const uint32_t *seq_id,
                                   uint32_t n_prot, uint32_t *bitmap, cudaStream_t stream) {
    if (n_prot == 0) return cudaSuccess;
    signature_flags_kernel<<<(n_prot + 255) / 256, 256, 0, stream>>>(prot_windows, prot_rejected, seq_id, n_prot, bitmap);
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
