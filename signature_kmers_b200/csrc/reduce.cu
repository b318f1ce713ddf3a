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
    const float thresh = __fmul_rn(__int2float_rn((int)count), 0.8f);      // tcc:250
    return !(__int2float_rn((int)best_count) < thresh);                    // tcc:254
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

// ---- the fused streaming reduce ---------------------------------------------
constexpr int RED_THREADS = 256;
constexpr int RED_WARPS = RED_THREADS / 32;
constexpr int RED_CHUNK = REDUCE_TILE / RED_WARPS;      // records per warp

struct RedSmem {
    uint64_t code[RED_WARPS][RED_CHUNK];
    uint32_t start[RED_WARPS][RED_CHUNK];
    uint32_t count[RED_WARPS][RED_CHUNK];               // group size; bit 31: needs order statistics
    uint16_t avg[RED_WARPS][RED_CHUNK];
    uint16_t func[RED_WARPS][RED_CHUNK];
    uint16_t mean[RED_WARPS][RED_CHUNK];
    uint32_t warp_rows[RED_WARPS];
    uint32_t warp_work[RED_WARPS];
    uint32_t warp_segs[RED_WARPS];
    uint32_t tile;
    uint32_t work_base;
    uint64_t base;
};

__global__ void __launch_bounds__(RED_THREADS)
fused_reduce_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals,
                    const uint64_t *__restrict__ n_ptr, const uint4 *__restrict__ meta,
                    const uint4 *__restrict__ giant_side, KeptColumns out, OrderWork *__restrict__ work,
                    uint32_t *__restrict__ n_work, uint32_t *__restrict__ bitmap, uint32_t *__restrict__ distinct_functions,
                    uint64_t *__restrict__ scan_state, uint32_t *__restrict__ ticket,
                    uint64_t *__restrict__ n_kept_out, uint64_t *__restrict__ n_seg_out, int order_stats) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RedSmem &sm = *reinterpret_cast<RedSmem *>(smem_raw);
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint64_t n = *n_ptr;
    if (tid == 0) sm.tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = sm.tile;
    const uint64_t tile_start = (uint64_t)tile * REDUCE_TILE;
    if (tile_start >= n) return;

    const uint64_t c0 = tile_start + (uint64_t)warp * RED_CHUNK;
    const uint64_t c1 = (c0 + RED_CHUNK < n) ? c0 + RED_CHUNK : n;
    uint32_t rows = 0, segs = 0, nwork = 0;

    // first group head at or after c0
    uint64_t cur = c1;
    for (uint64_t p0 = c0; p0 < c1; p0 += 32) {
        const uint64_t p = p0 + lane;
        bool head = false;
        if (p < c1) head = (p == 0) || (sigk_key_code(__ldg(keys + p)) != sigk_key_code(__ldg(keys + p - 1)));   // kmer != cur, tcc:194
        const unsigned hb = __ballot_sync(FULL, head);
        if (hb) { cur = p0 + __ffs(hb) - 1; break; }
    }

    while (cur < c1) {
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
            const int last = 31 - __clz(H);             // head of the group that runs past the window
            if (last == 0) {
                // ---- a group longer than 32 records: pre-reduced (giant) or walked now.
                // At most one long group can have its head in a 32-record block, so a
                // valid side entry under cur >> 5 is this group's.
                SegResult r;
                uint32_t glen;
                const uint4 g = __ldg(giant_side + (cur >> 5));
                if (g.x & 1u) {
                    glen = g.z;
                    r.keep = (g.x & 2u) != 0; r.func = g.y & 0xFFFFu; r.avg = g.y >> 16; r.mean = g.x >> 16; r.best_count = g.w;
                } else {
                    uint64_t e = cur + 32;              // the first 33 records are known to match
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
                const uint64_t code0 = __shfl_sync(FULL, code, 0);
                ++segs;
                if (r.keep && lane == 0) {
                    sm.code[warp][rows] = code0;
                    sm.start[warp][rows] = (uint32_t)cur;
                    sm.count[warp][rows] = glen | (r.best_count >= 2 ? 0x80000000u : 0u);
                    sm.avg[warp][rows] = (uint16_t)r.avg;
                    sm.func[warp][rows] = (uint16_t)r.func;
                    sm.mean[warp][rows] = (uint16_t)r.mean;
                    atomicAdd(distinct_functions + r.func, 1u);           // tcc:286
                }
                if (r.keep) { ++rows; if (r.best_count >= 2) ++nwork; }
                cur += glen;
                continue;
            }
            wlen = (uint32_t)last;                      // drop the unfinished group from this window
            H &= mask_lt((unsigned)last);
        }
        // groups headed at or beyond c1 belong to the next chunk
        if (cur + wlen > c1) {
            const unsigned beyond = H & ~mask_lt((unsigned)(c1 - cur));
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

        // function vote (func_count + arg-max, tcc:203, :228-248)
        uint32_t cand = __shfl_sync(FULL, f, s_lane);
        if (!__all_sync(FULL, !act || f == cand)) {
            uint32_t c = 0;
#pragma unroll
            for (int b = 0; b < 16; ++b) {
                const unsigned bal = __ballot_sync(FULL, act && ((f >> b) & 1u)) & segmask;
                c |= (2u * (uint32_t)__popc(bal) > cnt ? 1u : 0u) << b;
            }
            cand = c;
        }
        const bool is_best = act && f == cand;
        const unsigned best_mask = __ballot_sync(FULL, is_best) & segmask;
        const uint32_t best_count = __popc(best_mask);
        const bool keep = act && keep_rule(best_count, cnt);
        const unsigned K = __ballot_sync(FULL, keep && head);

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

        if (keep) mark_sequence(bitmap, m.y);
        if (keep && head) {
            const uint32_t slot = rows + __popc(K & mask_lt(lane));
            sm.code[warp][slot] = code;
            sm.start[warp][slot] = (uint32_t)p;
            sm.count[warp][slot] = cnt | (best_count >= 2 ? 0x80000000u : 0u);
            sm.avg[warp][slot] = (uint16_t)sel;
            sm.func[warp][slot] = (uint16_t)cand;
            sm.mean[warp][slot] = (uint16_t)(S / best_count);             // u16((double)S / n), exact
            atomicAdd(distinct_functions + cand, 1u);                      // tcc:286
        }
        rows += __popc(K);
        nwork += __popc(__ballot_sync(FULL, keep && head && best_count >= 2));
        segs += __popc(H);
        cur += wlen;
    }

    // ---- ordered write-out: warp totals -> tile base by chained scan -> rows
    if (lane == 0) { sm.warp_rows[warp] = rows; sm.warp_work[warp] = nwork; sm.warp_segs[warp] = segs; }
    __syncthreads();
    if (tid == 0) {
        uint32_t run = 0, sg = 0, wk = 0;
        for (int w = 0; w < RED_WARPS; ++w) {
            const uint32_t v = sm.warp_rows[w]; sm.warp_rows[w] = run; run += v;
            const uint32_t u = sm.warp_work[w]; sm.warp_work[w] = wk; wk += u;
            sg += sm.warp_segs[w];
        }
        const uint64_t excl = chained_scan_exclusive(scan_state, tile, run);
        sm.base = excl;
        sm.work_base = (order_stats && wk) ? atomicAdd(n_work, wk) : 0u;
        if (sg) atomicAdd(reinterpret_cast<unsigned long long *>(n_seg_out), (unsigned long long)sg);
        if (tile_start + REDUCE_TILE >= n) *n_kept_out = excl + run;
    }
    __syncthreads();
    const uint64_t obase = sm.base + sm.warp_rows[warp];
    uint32_t wslot = sm.work_base + sm.warp_work[warp];
    for (uint32_t i0 = 0; i0 < rows; i0 += 32) {
        const uint32_t i = i0 + lane;
        const bool have = i < rows;
        uint32_t cw = 0;
        if (have) {
            const uint64_t o = obase + i;
            cw = sm.count[warp][i];
            out.kmer[o] = sigk_code_to_ascii(sm.code[warp][i]);
            out.avg_from_end[o] = sm.avg[warp][i];
            out.function_index[o] = sm.func[warp][i];
            out.mean[o] = sm.mean[warp][i];
            out.median[o] = 0;
            out.var[o] = 0;
        }
        const bool need = have && order_stats && (cw & 0x80000000u);
        const unsigned nb = __ballot_sync(FULL, need);
        if (need) work[wslot + __popc(nb & mask_lt(lane))] = OrderWork{(uint32_t)(obase + i), sm.start[warp][i], cw & 0x7FFFFFFFu};
        wslot += __popc(nb);
    }
}

// ---- order-dependent columns: median (P^2) and var (iterative), one thread per kept group
__global__ void __launch_bounds__(128)
order_stats_kernel(const uint32_t *__restrict__ vals, const uint4 *__restrict__ meta, const OrderWork *__restrict__ work,
                   const uint32_t *__restrict__ n_work, KeptColumns out) {
    const uint32_t total = *n_work;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const OrderWork w = work[i];
        const uint32_t cand = out.function_index[w.row];
        LengthAcc acc;
        // newest first: the multimap iterates a key's items in reverse insertion order
        for (uint32_t j = w.count; j-- > 0;) {
            const uint4 m = __ldg(meta + vals[(uint64_t)w.start + j]);
            if (m.z == cand) acc.push(m.x);                               // acc(item.protein_length), tcc:271
        }
        out.median[w.row] = (uint16_t)u16_from_double(acc.q2);            // tcc:278
        out.var[w.row] = (uint16_t)u16_from_double(acc.var);              // tcc:279
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
size_t reduce_work_entries(uint64_t capacity) { return (size_t)(capacity / 2 + 2); }

cudaError_t reduce_configure() {
    return cudaFuncSetAttribute(fused_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RedSmem));
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

cudaError_t launch_fused_reduce(const uint64_t *keys, const uint32_t *vals, const uint64_t *n_ptr, uint64_t capacity,
                                const uint4 *meta, const uint4 *giant_side, KeptColumns out, OrderWork *work,
                                uint32_t *n_work, uint32_t *bitmap, uint32_t *distinct_functions, uint64_t *scan_state,
                                uint32_t *ticket, uint64_t *n_kept_out, uint64_t *n_seg_out, int order_stats,
                                cudaStream_t stream) {
    if (capacity == 0) return cudaSuccess;
    fused_reduce_kernel<<<(unsigned)reduce_tiles(capacity), RED_THREADS, sizeof(RedSmem), stream>>>(
        keys, vals, n_ptr, meta, giant_side, out, work, n_work, bitmap, distinct_functions, scan_state, ticket, n_kept_out,
        n_seg_out, order_stats);
    return cudaGetLastError();
}

cudaError_t launch_order_stats(const uint32_t *vals, const uint4 *meta, const OrderWork *work, const uint32_t *n_work,
                               uint64_t capacity, KeptColumns out, int sm_count, cudaStream_t stream) {
    if (capacity == 0) return cudaSuccess;
    order_stats_kernel<<<sm_count * 16, 128, 0, stream>>>(vals, meta, work, n_work, out);
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
