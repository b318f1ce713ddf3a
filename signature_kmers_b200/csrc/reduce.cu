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
//   statistics            :274, :285-286   -> seq bitmap, per-function counts
//
// Majority vote instead of a tally: a group is kept only if best_count >=
// 0.8*count, so a kept group's best function is a strict majority.  For a strict
// majority element every bit of its index is the majority value of that bit, so
// 16 ballot+popc steps give the only possible candidate; counting it decides.
// If it is not a strict majority no function is, the true best_count is
// <= count/2 and the float test rejects for every count — ties never reach a
// kept row, so the std::map walk's "lowest index wins" cannot matter.
//
// Work mapping (fused_reduce_kernel): a warp streams its chunk of the sorted
// records in windows of <= 32 records that start at a group head and contain
// only whole groups; lane = record, and every group in the window is reduced at
// once with ballots and segmented shuffles.  Groups longer than 32 records are
// walked 32 at a time by the whole warp; groups long enough to stall a tile
// (>= 513 records) are found by sampling and reduced by a pre-pass.  Rows are
// staged in shared memory and written in k-mer order through a chained scan of
// kept counts over the tiles.
//
// HBM traffic per record: 12 B read (+ a 16-byte L2 gather of the protein's
// meta); per kept row 18 B written.
#include "kernels.h"
#include "sigk_common.cuh"
#include "length_acc.cuh"

namespace sigk {

namespace {

constexpr unsigned FULL = 0xffffffffu;

SIGK_D unsigned mask_lt(unsigned lane) { return (1u << lane) - 1u; }
SIGK_D unsigned mask_le(unsigned lane) { return (2u << lane) - 1u; }      // lane 31 -> 0xffffffff (shift wraps to 0, -1)
SIGK_D unsigned mask_range(unsigned lo, unsigned hi) {                     // bits [lo, hi), hi <= 32
    const unsigned upper = hi >= 32 ? FULL : ((1u << hi) - 1u);
    return upper & ~((1u << lo) - 1u);
}

SIGK_D bool keep_rule(uint32_t best_count, uint32_t count) {
    if (2ull * best_count <= count) return false;                          // no strict majority (see header)
    // reject iff (float)best_count < float(count) * 0.8f (tcc:250-257).  0.8*count is at least 0.2
    // away from any integer it does not equal and the float product is off by < count * 6e-8, so
    // below 2^20 the float test is exactly 5*best >= 4*count; above, run the float ops themselves.
    if (count < (1u << 20)) return 5u * best_count >= 4u * count;
    const float thresh = __fmul_rn(__int2float_rn((int)count), 0.8f);
    return !(__int2float_rn((int)best_count) < thresh);
}

SIGK_D void mark_sequence(uint32_t *bitmap, uint32_t sid) {                // seqs_with_a_signature.insert, tcc:274
    const uint32_t bit = 1u << (sid & 31u);
    uint32_t *w = bitmap + (sid >> 5);
    if (!(__ldcg(w) & bit)) atomicOr(w, bit);
}

struct SegResult {
    bool keep;
    uint32_t func, best_count, avg, mean;
};

// A group of n > 32 records starting at `start`, reduced by the whole warp in
// strides of 32 (six walks over the group: vote, count+sum, four select rounds).
SIGK_D SegResult reduce_long_segment(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals,
                                     const uint4 *__restrict__ meta, uint64_t start, uint32_t n, uint32_t *bitmap) {
    const unsigned lane = threadIdx.x & 31u;
    SegResult r{false, 0, 0, 0, 0};
    // walk 1: bit-sliced majority vote over func_index; lane b owns bit b
    uint32_t ones = 0;
    for (uint32_t base = 0; base < n; base += 32) {
        const uint32_t j = base + lane;
        const bool act = j < n;
        const uint32_t f = act ? __ldg(&meta[vals[start + j]]).z : 0u;
#pragma unroll
        for (int b = 0; b < 16; ++b) {
            const unsigned bal = __ballot_sync(FULL, act && ((f >> b) & 1u));
            if ((int)lane == b) ones += __popc(bal);
        }
    }
    const uint32_t cand = __ballot_sync(FULL, lane < 16 && 2ull * ones > n) & 0xFFFFu;
    // walk 2: count the candidate, sum its lengths (mod 65536)
    uint32_t best = 0, S = 0;
    for (uint32_t base = 0; base < n; base += 32) {
        const uint32_t j = base + lane;
        if (j < n) {
            const uint4 m = __ldg(&meta[vals[start + j]]);
            if (m.z == cand) { ++best; S += m.x; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { best += __shfl_xor_sync(FULL, best, o); S += __shfl_xor_sync(FULL, S, o); }
    r.func = cand;
    r.best_count = best;
    r.keep = keep_rule(best, n);
    if (!r.keep) return r;
    r.mean = (S & 0xFFFFu) / best;
    // walks 3-6: offset of rank n/2 by radix select, 4 bits per round; lane v owns nibble value v
    uint32_t rank = n / 2, prefix = 0, pmask = 0;
#pragma unroll 1
    for (int shift = 12; shift >= 0; shift -= 4) {
        uint32_t cnt = 0;
        for (uint32_t base = 0; base < n; base += 32) {
            const uint32_t j = base + lane;
            const bool act = j < n;
            uint32_t off = 0;
            if (act) {
                off = sigk_key_offset(keys[start + j]);
                if (shift == 12) mark_sequence(bitmap, __ldg(&meta[vals[start + j]]).y);
            }
            const bool match = act && ((off & pmask) == prefix);
            const uint32_t nib = (off >> shift) & 15u;
#pragma unroll
            for (int v = 0; v < 16; ++v) {
                const unsigned bal = __ballot_sync(FULL, match && nib == (uint32_t)v);
                if ((int)lane == v) cnt += __popc(bal);
            }
        }
        uint32_t incl = lane < 16 ? cnt : 0u;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL, incl, o);
            if (lane >= (unsigned)o) incl += y;
        }
        const unsigned over = __ballot_sync(FULL, lane < 16 && incl > rank);
        const int sel = __ffs(over) - 1;
        const uint32_t below = __shfl_sync(FULL, incl - (lane < 16 ? cnt : 0u), sel);
        rank -= below;
        prefix |= (uint32_t)sel << shift;
        pmask |= 15u << shift;
    }
    r.avg = prefix;
    return r;
}

// ---- giant groups: found by sampling every GIANT_STRIDE records -------------
constexpr int GIANT_STRIDE = 512;

struct GiantEntry { uint32_t start, n; };

// giant_side[start >> 5]: x = 1 (valid) | keep << 1 | mean << 16, y = func | avg << 16, z = n, w = best_count
__global__ void giant_find_kernel(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ n_ptr,
                                  GiantEntry *__restrict__ list, uint32_t *__restrict__ n_list) {
    const uint64_t n = *n_ptr;
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t q = j * GIANT_STRIDE;
    if (q + GIANT_STRIDE >= n) return;
    const uint64_t c = sigk_key_code(keys[q]);
    if (sigk_key_code(keys[q + GIANT_STRIDE]) != c) return;              // does not span two samples
    if (j > 0 && sigk_key_code(keys[q - GIANT_STRIDE]) == c) return;     // an earlier sample owns it
    // head: first p in (q - STRIDE, q] with code == c (codes are sorted)
    uint64_t lo = j > 0 ? q - GIANT_STRIDE + 1 : 0, hi = q;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (sigk_key_code(keys[mid]) >= c) hi = mid; else lo = mid + 1;
    }
    const uint64_t start = lo;
    // end: gallop, then bisect, for the first e with code != c
    uint64_t step = GIANT_STRIDE, a = q + GIANT_STRIDE, b;
    for (;;) {
        b = a + step;
        if (b >= n) { b = n; break; }
        if (sigk_key_code(keys[b]) != c) break;
        a = b;
        step <<= 1;
    }
    // invariant: code[a] == c, (b == n or code[b] != c)
    while (a + 1 < b) {
        const uint64_t mid = (a + b) >> 1;
        if (sigk_key_code(keys[mid]) == c) a = mid; else b = mid;
    }
    const uint32_t slot = atomicAdd(n_list, 1u);
    list[slot] = GiantEntry{(uint32_t)start, (uint32_t)(b - start)};
}

__global__ void __launch_bounds__(256)
giant_reduce_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, const uint4 *__restrict__ meta,
                    const GiantEntry *__restrict__ list, const uint32_t *__restrict__ n_list, uint32_t *__restrict__ next,
                    uint4 *__restrict__ giant_side, uint32_t *__restrict__ bitmap) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t total = *n_list;
    for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = atomicAdd(next, 1u);
        i = __shfl_sync(FULL, i, 0);
        if (i >= total) return;
        const GiantEntry g = list[i];
        const SegResult r = reduce_long_segment(keys, vals, meta, g.start, g.n, bitmap);
        if (lane == 0)
            giant_side[g.start >> 5] = make_uint4(1u | (r.keep ? 2u : 0u) | (r.mean << 16), r.func | (r.avg << 16), g.n, r.best_count);
    }
}

// ---- the streaming reduce ----------------------------------------------------
// Persistent warps.  A warp takes a batch of RED_BATCH sorted records by ticket,
// counts the group heads in it and publishes that count at once (chained scan
// over batches), so later batches never wait for this one's work: every group,
// kept or not, owns the row slot of its index, and rejected groups leave a
// tombstone (function_index 0xFFFF) that squeeze_rows_kernel removes.
//
// rows[g] (uint4): x = code[31:0]; y = code[42:32] | avg_from_end << 11;
//                  z = function_index | mean << 16; w = median | var << 16
constexpr int RED_THREADS = 128;
constexpr int WORK_BLOCK = 64;          // order-statistics work slots a warp reserves at a time
constexpr uint32_t ORD_LONG = 4096;     // groups above this are walked by a whole warp (the tail); below, a lane each

struct WorkCursor { uint32_t base, free; };

SIGK_D void work_reserve(WorkCursor &wc, uint32_t need, OrderWork *__restrict__ work, uint32_t *__restrict__ n_work) {
    const unsigned lane = threadIdx.x & 31u;
    if (wc.free >= need) return;
    for (uint32_t i = lane; i < wc.free; i += 32) work[wc.base + i] = OrderWork{0u, 0u, 0u};   // unused tail: count 0 = skip
    uint32_t b = 0;
    if (lane == 0) b = atomicAdd(n_work, (uint32_t)WORK_BLOCK);
    wc.base = __shfl_sync(FULL, b, 0);
    wc.free = WORK_BLOCK;
}

__global__ void __launch_bounds__(RED_THREADS)
stream_reduce_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals,
                     const uint64_t *__restrict__ n_ptr, const uint4 *__restrict__ meta,
                     const uint4 *__restrict__ giant_side, uint4 *__restrict__ rows, OrderWork *__restrict__ work,
                     uint32_t *__restrict__ n_work, OrderWork *__restrict__ work_long, uint32_t *__restrict__ n_work_long,
                     uint32_t *__restrict__ bitmap, uint32_t *__restrict__ distinct_functions,
                     uint64_t *__restrict__ scan_state, uint32_t *__restrict__ ticket, uint64_t *__restrict__ n_seg_out,
                     int order_stats) {
    const unsigned lane = threadIdx.x & 31u;
    const uint64_t n = *n_ptr;
    WorkCursor wc{0u, 0u};

    for (;;) {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(ticket, 1u);
        t = __shfl_sync(FULL, t, 0);
        const uint64_t b0 = (uint64_t)t * RED_BATCH;
        if (b0 >= n) break;
        const uint64_t b1 = (b0 + RED_BATCH < n) ? b0 + RED_BATCH : n;

        // ---- group heads of the batch: count, first head; publish the count
        uint32_t hc = 0;
        uint64_t cur = b1;
        for (uint64_t p0 = b0; p0 < b1; p0 += 32) {
            const uint64_t p = p0 + lane;
            bool head = false;
            if (p < b1) head = (p == 0) || (sigk_key_code(__ldg(keys + p)) != sigk_key_code(__ldg(keys + p - 1)));   // kmer != cur, tcc:194
            const unsigned hb = __ballot_sync(FULL, head);
            if (hb && cur == b1) cur = p0 + __ffs(hb) - 1;
            hc += __popc(hb);
        }
        uint64_t g = chained_scan_exclusive_warp(scan_state, t, hc);        // index of the batch's first group
        if (b1 == n && lane == 0) *n_seg_out = g + hc;

        while (cur < b1) {
            // ---- one window: records cur .. cur+31, cur is a group head
            const uint64_t p = cur + lane;
            const bool valid = p < n;
            const uint64_t key = valid ? __ldg(keys + p) : 0ull;
            const uint64_t code = sigk_key_code(key);
            const uint64_t prev = __shfl_up_sync(FULL, code, 1);
            const bool head = valid && (lane == 0 || code != prev);
            unsigned H = __ballot_sync(FULL, head);
            const unsigned V = __ballot_sync(FULL, valid);
            // is the record after the window a head (or the end)?
            uint64_t nxt = 0;
            if (lane == 31) nxt = (cur + 32 < n) ? sigk_key_code(__ldg(keys + cur + 32)) : ~0ull;
            const bool closed = __shfl_sync(FULL, (lane == 31) && (!valid || nxt != code), 31);
            uint32_t wlen = __popc(V);
            if (!closed) {
                const int last = 31 - __clz(H);         // head of the group that runs past the window
                if (last == 0) {
                    // ---- a group longer than 32 records: pre-reduced (giant) or walked now.
                    // At most one long group can have its head in a 32-record block, so a
                    // valid side entry under cur >> 5 is this group's.
                    SegResult r;
                    uint32_t glen;
                    const uint4 gs = __ldg(giant_side + (cur >> 5));
                    if (gs.x & 1u) {
                        glen = gs.z;
                        r.keep = (gs.x & 2u) != 0; r.func = gs.y & 0xFFFFu; r.avg = gs.y >> 16; r.mean = gs.x >> 16; r.best_count = gs.w;
                    } else {
                        uint64_t e = cur + 32;          // the first 33 records are known to match
                        for (;;) {
                            const uint64_t q = e + lane;
                            const bool same = q < n && sigk_key_code(__ldg(keys + q)) == code;
                            const unsigned sb = __ballot_sync(FULL, same);
                            if (sb != FULL) { e += (uint64_t)(__ffs(~sb) - 1); break; }
                            e += 32;
                        }
                        glen = (uint32_t)(e - cur);
                        r = reduce_long_segment(keys, vals, meta, cur, glen, bitmap);
                    }
                    const bool walk = r.keep && order_stats;          // best_count >= 27 here
                    const bool walk_long = walk && glen > ORD_LONG;   // whole-warp walk (order_stats_long_kernel)
                    if (walk && !walk_long) work_reserve(wc, 1u, work, n_work);
                    if (lane == 0) {
                        if (r.keep) {
                            rows[g] = make_uint4((uint32_t)code, (uint32_t)(code >> 32) | (r.avg << 11), r.func | (r.mean << 16), 0u);
                            atomicAdd(distinct_functions + r.func, 1u);                 // tcc:286
                            if (walk_long) work_long[atomicAdd(n_work_long, 1u)] = OrderWork{(uint32_t)g, (uint32_t)cur, glen};
                            else if (walk) work[wc.base] = OrderWork{(uint32_t)g, (uint32_t)cur, glen};
                        } else rows[g] = make_uint4(0u, 0u, 0xFFFFu, 0u);
                    }
                    if (walk && !walk_long) { wc.base += 1; wc.free -= 1; }
                    g += 1;
                    cur += glen;
                    continue;
                }
                wlen = (uint32_t)last;                  // drop the unfinished group from this window
                H &= mask_lt((unsigned)last);
            }
            // groups headed at or beyond b1 belong to the next batch
            if (cur + wlen > b1) {
                const unsigned beyond = H & ~mask_lt((unsigned)(b1 - cur));
                if (beyond) { wlen = (uint32_t)(__ffs(beyond) - 1); H &= mask_lt(wlen); }
            }
            const bool act = lane < wlen;
            const uint32_t ord = act ? __ldg(vals + p) : 0u;
            const uint4 m = act ? __ldg(meta + ord) : make_uint4(0, 0, 0, 0);      // len, seq_id, func
            const uint32_t f = m.z;
            const uint32_t off = sigk_key_offset(key);

            // my group's lanes
            const unsigned s_lane = 31u - (unsigned)__clz(H & mask_le(lane));
            const unsigned above = H & ~mask_le(lane);
            const unsigned e_lane = above ? (unsigned)(__ffs(above) - 1) : wlen;
            const uint32_t cnt = e_lane - s_lane;
            const unsigned segmask = mask_range(s_lane, e_lane);

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
            if (keep) mark_sequence(bitmap, m.y);

            const bool walk = keep && head && order_stats && best_count >= 3;
            const unsigned wb = __ballot_sync(FULL, walk);
            if (wb) work_reserve(wc, (uint32_t)__popc(wb), work, n_work);
            if (head && act) {
                const uint64_t gi = g + __popc(H & mask_lt(lane));
                if (keep) {
                    // u16((double)S / n) = floor(S / best_count), S < 65536, best_count <= 32: float quotient, fixed up
                    uint32_t mean = (uint32_t)__float2uint_rz(__fmul_rn((float)S, __frcp_rn((float)best_count)));
                    if (mean * best_count > S) --mean;
                    else if ((mean + 1u) * best_count <= S) ++mean;
                    rows[gi] = make_uint4((uint32_t)code, (uint32_t)(code >> 32) | (sel << 11), cand | (mean << 16), var2 << 16);
                    atomicAdd(distinct_functions + cand, 1u);                                   // tcc:286
                    if (walk) work[wc.base + __popc(wb & mask_lt(lane))] = OrderWork{(uint32_t)gi, (uint32_t)p, cnt};
                } else rows[gi] = make_uint4(0u, 0u, 0xFFFFu, 0u);
            }
            if (wb) { const uint32_t k = __popc(wb); wc.base += k; wc.free -= k; }
            g += __popc(H);
            cur += wlen;
        }
    }
    // the reserved slots this warp never used
    for (uint32_t i = lane; i < wc.free; i += 32) work[wc.base + i] = OrderWork{0u, 0u, 0u};
}

// ---- squeeze: drop the tombstones, keep k-mer order, expand rows into the table columns
constexpr int SQ_THREADS = 256;
constexpr int SQ_ITEMS = 8;
constexpr int SQ_TILE = SQ_THREADS * SQ_ITEMS;
constexpr int SQ_WARPS = SQ_THREADS / 32;

// two residues at a time: pair_ascii[40 a + b] = letter(a) | letter(b) << 8
__device__ uint16_t g_pair_ascii[1600];
__global__ void init_pair_ascii_kernel() {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 1600) g_pair_ascii[i] = (uint16_t)(sigk_symbol_ascii(i / 40) | (sigk_symbol_ascii(i % 40) << 8));
}
SIGK_D uint64_t code_to_ascii_pairs(uint64_t code) {
    const uint32_t hi = (uint32_t)(code / 2560000ull);                      // 40^4
    const uint32_t lo = (uint32_t)(code - (uint64_t)hi * 2560000ull);
    const uint32_t w0 = __ldg(g_pair_ascii + hi / 1600u) | ((uint32_t)__ldg(g_pair_ascii + hi % 1600u) << 16);
    const uint32_t w1 = __ldg(g_pair_ascii + lo / 1600u) | ((uint32_t)__ldg(g_pair_ascii + lo % 1600u) << 16);
    return (uint64_t)w0 | ((uint64_t)w1 << 32);
}

__global__ void __launch_bounds__(SQ_THREADS)
squeeze_rows_kernel(const uint4 *__restrict__ rows, const uint64_t *__restrict__ n_seg_ptr, KeptColumns out,
                    uint64_t *__restrict__ scan_state, uint32_t *__restrict__ ticket, uint64_t *__restrict__ n_kept_out) {
    __shared__ uint32_t s_scan[SQ_WARPS + 2];
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_base;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint64_t n_seg = *n_seg_ptr;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t tile_start = (uint64_t)tile * SQ_TILE;
    if (tile_start >= n_seg) return;

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
    if (tid == 0) chained_scan_publish(scan_state, tile, total);
    // the code -> ASCII expansion needs nothing from the other tiles: do it while they publish
    uint64_t ascii[SQ_ITEMS];
#pragma unroll
    for (int i = 0; i < SQ_ITEMS; ++i)
        ascii[i] = code_to_ascii_pairs((uint64_t)row[i].x | ((uint64_t)(row[i].y & 0x7FFu) << 32));
    if (tid == 0) {
        const uint64_t base = chained_scan_resolve(scan_state, tile, total);
        s_base = base;
        if (tile_start + SQ_TILE >= n_seg) *n_kept_out = base + total;
    }
    __syncthreads();
    const uint64_t base = s_base;
#pragma unroll
    for (int i = 0; i < SQ_ITEMS; ++i) {
        if ((ball[i] >> lane) & 1u) {
            const uint64_t o = base + run + __popc(ball[i] & mask_lt(lane));
            out.kmer[o] = ascii[i];
            out.avg_from_end[o] = (uint16_t)(row[i].y >> 11);
            out.function_index[o] = (uint16_t)(row[i].z & 0xFFFFu);
            out.mean[o] = (uint16_t)(row[i].z >> 16);
            out.median[o] = (uint16_t)(row[i].w & 0xFFFFu);
            out.var[o] = (uint16_t)(row[i].w >> 16);
        }
        run += __popc(ball[i]);
    }
}

// ---- order-dependent columns: median (P^2) and var (iterative) of the kept groups in `work`.
// The recurrences are sequential per group, so a lane owns one group at a time and a lane
// that finishes takes the next entry of its warp's block at once (groups differ in length by
// orders of magnitude; waiting for the slowest lane left 13 % of the lanes busy in the v2 profile).
constexpr int ORD_THREADS = 128;
constexpr int ORD_BLOCK = 256;      // work entries one warp owns at a time

__global__ void __launch_bounds__(ORD_THREADS)
order_stats_kernel(const uint32_t *__restrict__ vals, const uint4 *__restrict__ meta, const OrderWork *__restrict__ work,
                   const uint32_t *__restrict__ n_work, uint4 *__restrict__ rows) {
    const uint32_t total = *n_work;
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t warp_global = (blockIdx.x * ORD_THREADS + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * ORD_THREADS) >> 5;
    for (uint64_t b0 = (uint64_t)warp_global * ORD_BLOCK; b0 < total; b0 += (uint64_t)n_warps * ORD_BLOCK) {
        uint32_t cursor = (uint32_t)b0;
        const uint32_t b1 = (uint32_t)((b0 + ORD_BLOCK < total) ? b0 + ORD_BLOCK : total);
        bool active = false;
        uint32_t row = 0, cand = 0, left = 0;
        uint64_t start = 0;
        LengthAcc acc;
        for (;;) {
            const unsigned idle = __ballot_sync(FULL, !active);
            if (idle && cursor < b1) {
                const uint32_t take = min((uint32_t)__popc(idle), b1 - cursor);
                const uint32_t r = __popc(idle & mask_lt(lane));
                if (!active && r < take) {
                    const OrderWork w = work[cursor + r];
                    if (w.count) {                                         // count 0: an unused reserved slot
                        row = w.row; start = w.start; left = w.count;
                        cand = rows[w.row].z & 0xFFFFu;
                        acc = LengthAcc();
                        active = true;
                    }
                }
                cursor += take;
            }
            if (!__any_sync(FULL, active)) { if (cursor >= b1) break; else continue; }
            if (active) {
                // newest first: the multimap iterates a key's items in reverse insertion order
                --left;
                const uint4 m = __ldg(meta + __ldg(vals + start + left));
                if (m.z == cand) acc.push(m.x);                            // acc(item.protein_length), tcc:271
                if (left == 0) {
                    rows[row].w = u16_from_double(acc.q2) | (u16_from_double(acc.var) << 16);   // tcc:278-279
                    active = false;
                }
            }
        }
    }
}

// Groups of more than ORD_LONG records: one warp per group.  The recurrences stay sequential, but
// the warp fetches 32 records at a time (coalesced values, 32 meta gathers in flight, the next
// batch prefetched) and every lane runs the same accumulator on shuffled samples, so a group no
// longer pays two dependent memory latencies per record (524 ms -> see profiles/ on the Zipf set).
__global__ void __launch_bounds__(128)
order_stats_long_kernel(const uint32_t *__restrict__ vals, const uint4 *__restrict__ meta, const OrderWork *__restrict__ work,
                        const uint32_t *__restrict__ n_work, uint32_t *__restrict__ next, uint4 *__restrict__ rows) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t total = *n_work;
    for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = atomicAdd(next, 1u);
        i = __shfl_sync(FULL, i, 0);
        if (i >= total) return;
        const OrderWork w = work[i];
        const uint32_t cand = rows[w.row].z & 0xFFFFu;
        LengthAcc acc;
        // newest first: batch b covers records count-1-32b-lane
        int64_t j = (int64_t)w.count - 1 - (int64_t)lane;
        uint4 m = j >= 0 ? __ldg(meta + __ldg(vals + (uint64_t)w.start + j)) : make_uint4(0, 0, 0xFFFFFFFFu, 0);
        for (int64_t left = w.count; left > 0; left -= 32) {
            const int64_t jn = j - 32;
            const uint4 mn = (left > 32 && jn >= 0) ? __ldg(meta + __ldg(vals + (uint64_t)w.start + jn)) : make_uint4(0, 0, 0xFFFFFFFFu, 0);
            const int take = left < 32 ? (int)left : 32;
            for (int l = 0; l < take; ++l) {
                const uint32_t f = __shfl_sync(FULL, m.z, l);
                const uint32_t x = __shfl_sync(FULL, m.x, l);
                if (f == cand) acc.push(x);                                 // acc(item.protein_length), tcc:271
            }
            m = mn;
            j = jn;
        }
        if (lane == 0) rows[w.row].w = u16_from_double(acc.q2) | (u16_from_double(acc.var) << 16);
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

__global__ void protein_meta_kernel(const uint64_t *__restrict__ starts, const uint16_t *__restrict__ func,
                                    const uint32_t *__restrict__ seq_id, uint32_t n_prot, uint4 *__restrict__ meta,
                                    uint32_t *__restrict__ seqs_with_func) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_prot) return;
    // protein_length = static_cast<unsigned int>(seq.length()), tcc:178
    meta[i] = make_uint4((uint32_t)(starts[i + 1] - starts[i]), seq_id[i], func[i], 0u);
    if (seqs_with_func) atomicAdd(seqs_with_func + func[i], 1u);          // seqs_with_func[function_index]++, tcc:160
}

}  // namespace

size_t reduce_side_entries(uint64_t capacity) { return (size_t)(capacity / 32 + 2); }
size_t reduce_giant_entries(uint64_t capacity) { return (size_t)(capacity / GIANT_STRIDE + 2); }
size_t reduce_long_work_entries(uint64_t capacity) { return (size_t)(capacity / ORD_LONG + 2); }
static int reduce_grid(int sm_count) { return sm_count * 12; }
size_t reduce_work_entries(uint64_t capacity, int sm_count) {
    // a walked group has >= 3 records; every warp of the persistent grid can strand one reserved block
    return (size_t)(capacity / 3 + 2) + (size_t)reduce_grid(sm_count) * (RED_THREADS / 32) * WORK_BLOCK;
}

cudaError_t reduce_configure() {
    init_pair_ascii_kernel<<<(1600 + 255) / 256, 256>>>();
    return cudaGetLastError();
}

cudaError_t launch_protein_meta(const uint64_t *starts, const uint16_t *func, const uint32_t *seq_id, uint32_t n_prot,
                                uint4 *meta, uint32_t *seqs_with_func, cudaStream_t stream) {
    if (n_prot == 0) return cudaSuccess;
    protein_meta_kernel<<<(n_prot + 255) / 256, 256, 0, stream>>>(starts, func, seq_id, n_prot, meta, seqs_with_func);
    return cudaGetLastError();
}

cudaError_t launch_giant_prepass(const uint64_t *keys, const uint32_t *vals, const uint64_t *n_ptr, uint64_t capacity,
                                 const uint4 *meta, void *giant_list, uint32_t *n_giant, uint32_t *next_giant,
                                 uint4 *giant_side, uint32_t *bitmap, int sm_count, cudaStream_t stream) {
    if (capacity <= GIANT_STRIDE) return cudaSuccess;
    const uint64_t samples = capacity / GIANT_STRIDE + 1;
    giant_find_kernel<<<(unsigned)((samples + 255) / 256), 256, 0, stream>>>(keys, n_ptr, (GiantEntry *)giant_list, n_giant);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    giant_reduce_kernel<<<sm_count * 2, 256, 0, stream>>>(keys, vals, meta, (const GiantEntry *)giant_list, n_giant, next_giant,
                                                          giant_side, bitmap);
    return cudaGetLastError();
}

cudaError_t launch_stream_reduce(const uint64_t *keys, const uint32_t *vals, const uint64_t *n_ptr, uint64_t capacity,
                                 const uint4 *meta, const uint4 *giant_side, uint4 *rows, OrderWork *work, uint32_t *n_work,
                                 OrderWork *work_long, uint32_t *n_work_long,
                                 uint32_t *bitmap, uint32_t *distinct_functions, uint64_t *scan_state, uint32_t *ticket,
                                 uint64_t *n_seg_out, int order_stats, int sm_count, cudaStream_t stream) {
    if (capacity == 0) return cudaSuccess;
    uint64_t grid = reduce_grid(sm_count);
    const uint64_t want = (reduce_batches(capacity) + RED_THREADS / 32 - 1) / (RED_THREADS / 32);
    if (grid > want) grid = want;
    stream_reduce_kernel<<<(unsigned)grid, RED_THREADS, 0, stream>>>(keys, vals, n_ptr, meta, giant_side, rows, work, n_work,
                                                                     work_long, n_work_long, bitmap, distinct_functions, scan_state, ticket, n_seg_out,
                                                                     order_stats);
    return cudaGetLastError();
}

cudaError_t launch_order_stats(const uint32_t *vals, const uint4 *meta, const OrderWork *work, const uint32_t *n_work,
                               const OrderWork *work_long, const uint32_t *n_work_long, uint32_t *next_long,
                               uint64_t capacity, uint4 *rows, int sm_count, cudaStream_t stream) {
    if (capacity == 0) return cudaSuccess;
    // the long groups first: they are the tail
    order_stats_long_kernel<<<sm_count * 8, 128, 0, stream>>>(vals, meta, work_long, n_work_long, next_long, rows);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    order_stats_kernel<<<sm_count * 16, ORD_THREADS, 0, stream>>>(vals, meta, work, n_work, rows);
    return cudaGetLastError();
}

cudaError_t launch_squeeze_rows(const uint4 *rows, const uint64_t *n_seg_ptr, uint64_t capacity, KeptColumns out,
                                uint64_t *scan_state, uint32_t *ticket, uint64_t *n_kept_out, cudaStream_t stream) {
    if (capacity == 0) return cudaSuccess;
    squeeze_rows_kernel<<<(unsigned)squeeze_tiles(capacity), SQ_THREADS, 0, stream>>>(rows, n_seg_ptr, out, scan_state, ticket, n_kept_out);
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
