// onesweep.cu — stage 2: stable least-significant-digit radix sort of the
// (key u64, value u32) records, onesweep style, with the encode fused into the first pass.
//
//   - every pass is ONE kernel of persistent CTAs: a CTA takes a tile by ticket, ranks its keys with
//     warp ballots (stable; one shared-memory atomic per digit group and warp), publishes its per-digit
//     counts, reorders the tile in shared memory, resolves its global offsets by decoupled look-back
//     over the previous tiles (by then they have usually published) and writes each digit's run with
//     coalesced stores;
//   - the digit counts of all passes exist before the first record does: window_count_kernel
//     (encode.cu) takes them from the residues;
//   - encode_sort_kernel is the window loop (window_scan.cuh) feeding that machinery directly: the
//     records of 14 slices are compacted in shared memory in canonical order and leave the CTA already
//     sorted on the lowest digit — the canonical-order record stream never exists in HBM;
//   - the first pass also splits the two runs: records whose window holds a lower-case residue
//     (mask8 != 0, a fraction of a per cent) go, in input order, to a side bin behind the main
//     records.  The main run is then sorted on its 35 code bits (3 more passes), the side run on
//     all 43 bits (its own histogram + 5 passes over a tiny array).
//
// This replaces the grouping the reference gets from hashing every occurrence
// into tbb::concurrent_unordered_multimap (src/signature_build.tcc:178) and
// walking the buckets (:186-208): after the last pass equal k-mers are adjacent
// and, because every pass is stable, still in insertion order.
//
// HBM traffic per record: fused first pass 1 B read + 12 B written; 24 B per later pass.
#include "kernels.h"
#include "sigk_common.cuh"
#include "window_scan.cuh"

#include <algorithm>

namespace sigk {

static_assert(SORT_RADIX_BITS == SIGK_RADIX_BITS && SORT_BINS == SIGK_BINS, "kernels.h and sigk_common.cuh agree");

PassPlan make_pass_plan(int bit_lo, int bit_hi) {
    PassPlan p{};
    const int total = bit_hi - bit_lo;
    int npass = (total + SIGK_RADIX_BITS - 1) / SIGK_RADIX_BITS;
    if (npass < 1) npass = 1;
    p.npass = npass;
    int lo = bit_lo;
    for (int i = 0; i < npass; ++i) {
        // spread the bits evenly: the first (total % npass) passes get one more
        const int bits = total / npass + (i < total % npass ? 1 : 0);
        p.lo[i] = lo;
        p.bits[i] = bits;
        lo += bits;
    }
    return p;
}

namespace {

constexpr int OS_WARPS = OS_THREADS / 32;
constexpr int OS_DPT = (SIGK_RADIX + OS_THREADS - 1) / OS_THREADS;      // digits one thread owns in the scan / look-back
static_assert(OS_DPT == 1 || OS_DPT == 2, "a thread owns one 16-bit counter per warp, or one 32-bit word of two");
static_assert(OS_THREADS >= 256, "the first 256 threads fill the symbol table");
static_assert(OS_TILE < 65536, "tile slots are kept in 16 bits");
constexpr uint32_t PAD_BIN = SIGK_SIDE_BIN + 1;      // rows of a short tile that hold no record
constexpr uint64_t PAD_KEY = ~0ull;                  // never a record: the low five key bits are zero

// tuning knobs (tools/sweep_variants.sh builds and times the alternatives)
#ifndef SIGK_OS_LATE_LOOKBACK
#define SIGK_OS_LATE_LOOKBACK 0                      // 1: resolve the look-back after the shared-memory reorder instead of before it
#endif
#ifndef SIGK_OS_BACKOFF_NS
#define SIGK_OS_BACKOFF_NS 0                         // __nanosleep between polls of a look-back word that is not ready
#endif
#ifndef SIGK_OS_LBW
#define SIGK_OS_LBW 1                                // predecessors whose look-back words are requested per round trip
#endif
#ifndef SIGK_OS_PIPELINE
#define SIGK_OS_PIPELINE 0                           // 1: request the next tile's keys before the write-out (measured: 11 ms per pass in situ against 4.4 ms — off)
#endif
#ifndef SIGK_OS_PERSISTENT
#define SIGK_OS_PERSISTENT 1                         // 1: a resident grid loops over tickets; 0: one CTA per tile
#endif

// look-back word: flag in the two top bits, value below
template <typename LB> struct LBTraits;
template <> struct LBTraits<uint32_t> {
    static constexpr uint32_t AGG = 1u << 30, PRE = 2u << 30, VAL = (1u << 30) - 1;
    static constexpr int SHIFT = 30;
    static SIGK_D uint32_t ld(const uint32_t *p) { return ld_volatile_u32(p); }
    static SIGK_D void st(uint32_t *p, uint32_t v) { st_volatile_u32(p, v); }
};
template <> struct LBTraits<uint64_t> {
    static constexpr uint64_t AGG = 1ull << 62, PRE = 2ull << 62, VAL = (1ull << 62) - 1;
    static constexpr int SHIFT = 62;
    static SIGK_D uint64_t ld(const uint64_t *p) { return ld_volatile_u64(p); }
    static SIGK_D void st(uint64_t *p, uint64_t v) { st_volatile_u64(p, v); }
};

// SLOTS: records the key / value staging holds (the fused kernel pads its canonical-order staging)
template <int SLOTS>
struct OsSmem {
    uint64_t keys[SLOTS];
    uint32_t vals[SLOTS];
    uint32_t goff[SIGK_BINS];               // global position of sorted slot 0 of each digit, minus its tile base (mod 2^32)
    // per-warp digit counters, two 16-bit counters to a word; after the scan: tile slot of the warp's first record of each digit
    uint32_t cnt[OS_WARPS][SIGK_BINS / 2];
    uint32_t scan[OS_WARPS + 2];
    uint32_t wtotal[OS_WARPS];
    uint32_t tile, next_tile;
    int8_t sym[256];
};
using PassSmem = OsSmem<OS_TILE>;
static_assert(ES_TILE == ES_WARPS * WS_SUB, "a fused tile is a whole number of slices");
static_assert(ES_WARPS <= OS_WARPS, "one encoding warp per slice");
constexpr int ES_STAGE = ES_TILE + ES_TILE / 16;
using FusedSmem = OsSmem<ES_STAGE>;
SIGK_D uint32_t stage_slot(uint32_t o) { return o + (o >> 4); }

// The digit of a record.  FIRST (the pass that also splits the runs): side records and padding get bins of their own.
template <bool FIRST>
SIGK_D uint32_t digit_of(uint64_t key, int bit_lo, uint32_t digit_mask) {
    uint32_t d = (uint32_t)(key >> bit_lo) & digit_mask;
    if (FIRST) {
        if (sigk_key_mask(key)) d = SIGK_SIDE_BIN;
        if (key == PAD_KEY) d = PAD_BIN;
    }
    return d;
}

// peers &= lanes whose digit has the same bit: ballot, a per-lane all-ones / all-zeros word from the bit, one
// three-input logic op — three instructions per digit bit plus one R2P per seven bits (the C form
// `peers &= bit ? m : ~m` compiled to seven per bit)
SIGK_D void peers_step(unsigned &peers, uint32_t d, uint32_t bit_mask) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 m, x;\n\t"
                 "and.b32 m, %1, %2;\n\t"
                 "setp.ne.u32 p, m, 0;\n\t"
                 "vote.sync.ballot.b32 m, p, 0xffffffff;\n\t"
                 "selp.b32 x, 0, 0xffffffff, p;\n\t"
                 "lop3.b32 %0, %0, m, x, 0x60;\n\t}"                  // peers & (m ^ x)
                 : "+r"(peers) : "r"(d), "r"(bit_mask) : "memory");
}

template <typename SM>
SIGK_D void os_zero_counters(SM &sm) {
    uint32_t *c32 = &sm.cnt[0][0];
    for (uint32_t j = threadIdx.x; j < OS_WARPS * (SIGK_BINS / 2); j += OS_THREADS) c32[j] = 0;

}

// Everything after the keys of a tile sit in registers (warp-striped: item i of lane l of warp w is tile record
// w * 32 ITEMS + 32 i + l; rows past tile_n hold PAD_KEY): rank, publish, reorder, look back, write out.
// The counters must be zero and visible (a barrier after os_zero_counters).
// val_at(r) = value of tile record r.  VALS_STAGED: val_at reads the staging that the sorted values are about to
// overwrite (fused kernel), so everybody reads before anybody writes.
// before_write_out() runs in every thread after the last barrier before the write-out, when the key registers and the
// digit counters are dead: the persistent pass kernel zeroes the counters there and starts loading its next tile.
template <typename LB, bool FIRST, bool VALS_STAGED, int ITEMS, typename SM, typename ValFn, typename MidFn>
SIGK_D void os_sort_tile(SM &sm, uint64_t (&key)[ITEMS], uint32_t tile, uint32_t tile_n, int bit_lo, uint32_t digit_mask,
                         const uint64_t *__restrict__ bin_base, LB *__restrict__ lookback, ValFn &&val_at,
                         uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, uint32_t *__restrict__ next_ticket,
                         MidFn &&before_write_out) {
    using T = LBTraits<LB>;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    uint16_t *wcnt16 = reinterpret_cast<uint16_t *>(sm.cnt[warp]);
    constexpr int NBITS = FIRST ? SIGK_RADIX_BITS + 1 : SIGK_RADIX_BITS;

    // ---- rank inside the warp: lanes with my digit are found with one ballot per digit bit (MATCH.ANY runs on
    // the ADU pipe at ~64 cycles per warp instruction: 70 % pipe utilisation in the round-1 v1 profile); the
    // group's first lane bumps the warp's counter with one shared-memory atomic whose return value is the
    // group's base; everyone takes base + (peers below me).  Stable.  The atomics of a chunk of items are
    // issued back to back so that their latencies overlap.
    const bool owner = tid * OS_DPT < SIGK_RADIX;
    uint32_t rank2[(ITEMS + 1) / 2];                    // two 16-bit ranks (later: tile slots) to a register
#pragma unroll
    for (int i = 0; i < (ITEMS + 1) / 2; ++i) rank2[i] = 0;
#define SIGK_RANK_GET(i) ((rank2[(i) >> 1] >> (16 * ((i) & 1))) & 0xFFFFu)
#define SIGK_RANK_SET(i, v) rank2[(i) >> 1] = ((i) & 1) ? ((rank2[(i) >> 1] & 0xFFFFu) | ((uint32_t)(v) << 16)) : ((rank2[(i) >> 1] & 0xFFFF0000u) | (uint32_t)(v))
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const uint32_t d = digit_of<FIRST>(key[i], bit_lo, digit_mask);
        unsigned peers = 0xffffffffu;
#pragma unroll
        for (int b = 0; b < NBITS; ++b) peers_step(peers, d, 1u << b);
        // every lane of the group reads the warp's counter (a broadcast); the group's first lane moves it on
        const uint32_t below = (uint32_t)__popc(peers & lt);
        const uint32_t old = wcnt16[d];
        if (below == 0) wcnt16[d] = (uint16_t)(old + (uint32_t)__popc(peers));
        SIGK_RANK_SET(i, old + below);
        __syncwarp();                                   // the next item's readers see this store
    }
    __syncthreads();

    // ---- per digit: exclusive prefix over warps, tile count, publish.  Thread t owns the OS_DPT digits
    // t * OS_DPT ..: one 16-bit counter per warp with 512 threads, one 32-bit word of two with 256.  In a FIRST
    // pass thread 0 also owns the side bin and the padding bin (which is never published).
    uint32_t my_count[OS_DPT], my_sum = 0;
#pragma unroll
    for (int k = 0; k < OS_DPT; ++k) my_count[k] = 0;
    if (owner) {
#pragma unroll
        for (int w = 0; w < OS_WARPS; ++w) {
            if (OS_DPT == 1) my_count[0] += reinterpret_cast<const uint16_t *>(sm.cnt[w])[tid];
            else { const uint32_t c = sm.cnt[w][tid]; my_count[0] += c & 0xFFFFu; my_count[OS_DPT - 1] += c >> 16; }
        }
#pragma unroll
        for (int k = 0; k < OS_DPT; ++k) {
            T::st(lookback + (size_t)tile * SIGK_BINS + tid * OS_DPT + k, (tile == 0 ? T::PRE : T::AGG) | (LB)my_count[k]);
            my_sum += my_count[k];
        }
    }
    uint32_t side_count = 0;
    if (FIRST && tid == 0) {
#pragma unroll
        for (int w = 0; w < OS_WARPS; ++w) side_count += reinterpret_cast<const uint16_t *>(sm.cnt[w])[SIGK_SIDE_BIN];
        T::st(lookback + (size_t)tile * SIGK_BINS + SIGK_SIDE_BIN, (tile == 0 ? T::PRE : T::AGG) | (LB)side_count);
    }
    uint32_t total_main;
    const uint32_t dbase = block_exclusive_scan<OS_THREADS>(my_sum, sm.scan, &total_main);
    if (owner) {
        // fold the digit's tile base into the per-warp prefixes: slot = cnt[warp][d] + rank
        uint32_t run0 = dbase, run1 = dbase + my_count[0];
#pragma unroll
        for (int w = 0; w < OS_WARPS; ++w) {
            if (OS_DPT == 1) {
                uint16_t *c = reinterpret_cast<uint16_t *>(sm.cnt[w]) + tid;
                const uint32_t x = *c;
                *c = (uint16_t)run0;
                run0 += x;
            } else {
                const uint32_t c = sm.cnt[w][tid];
                sm.cnt[w][tid] = run0 | (run1 << 16);
                run0 += c & 0xFFFFu;
                run1 += c >> 16;
            }
        }
    }
    if (FIRST && tid == 0) {
        uint32_t run = total_main;                      // side records sit behind every main digit, padding behind them
#pragma unroll
        for (int b = 0; b < 2; ++b) {
#pragma unroll
            for (int w = 0; w < OS_WARPS; ++w) {
                uint16_t *c = reinterpret_cast<uint16_t *>(sm.cnt[w]) + SIGK_SIDE_BIN + b;
                const uint32_t x = *c;
                *c = (uint16_t)run;
                run += x;
            }
        }
    }
    // ---- look back over the tiles before this one: per digit, add aggregates until an inclusive prefix turns up
    // The words of SIGK_OS_LBW predecessors are requested together: a walk of h tiles costs about h / LBW round trips
    // to L2 instead of h (the profile of the one-at-a-time walk: ~30 hops per tile, a quarter of all stall samples).
    auto look_back_digit = [&](uint32_t d, uint32_t cnt, uint32_t base) {
        LB excl = 0;
        if (tile > 0) {
            int64_t t = (int64_t)tile - 1;
            for (;;) {
                LB v[SIGK_OS_LBW];
#pragma unroll
                for (int k = 0; k < SIGK_OS_LBW; ++k)
                    v[k] = t - k >= 0 ? T::ld(lookback + (size_t)(t - k) * SIGK_BINS + d) : T::PRE;      // before tile 0: prefix 0
                bool done = false;
                int adv = 0;
#pragma unroll
                for (int k = 0; k < SIGK_OS_LBW; ++k) {
                    const LB flag = v[k] >> T::SHIFT;
                    if (!done && adv == k && flag != 0) {                   // everything nearer has been added
                        excl += v[k] & T::VAL;
                        adv = k + 1;
                        done = flag == 2;
                    }
                }
                if (done) break;
                if (adv == 0 && SIGK_OS_BACKOFF_NS) __nanosleep(SIGK_OS_BACKOFF_NS);
                t -= adv;                                                   // go on from the first word that was not ready
            }
            T::st(lookback + (size_t)tile * SIGK_BINS + d, T::PRE | (excl + (LB)cnt));
        }
        sm.goff[d] = (uint32_t)(bin_base[d] + (uint64_t)excl) - base;
    };
    // Thread 0 of a FIRST pass owns digit 0 and the side bin: the two walks advance together (one after the other they
    // put a second walk's latency in front of the tile's barrier: 7.7 ms per pass against 4.4 ms).
    auto look_back_pair = [&](uint32_t d0, uint32_t cnt0, uint32_t base0, uint32_t d1, uint32_t cnt1, uint32_t base1) {
        LB e0 = 0, e1 = 0;
        if (tile > 0) {
            int64_t t0 = (int64_t)tile - 1, t1 = t0;
            bool f0 = false, f1 = false;
            while (!(f0 && f1)) {
                const LB v0 = f0 ? (LB)0 : T::ld(lookback + (size_t)t0 * SIGK_BINS + d0);
                const LB v1 = f1 ? (LB)0 : T::ld(lookback + (size_t)t1 * SIGK_BINS + d1);
                if (!f0 && (v0 >> T::SHIFT) != 0) { e0 += v0 & T::VAL; if ((v0 >> T::SHIFT) == 2) f0 = true; else --t0; }
                if (!f1 && (v1 >> T::SHIFT) != 0) { e1 += v1 & T::VAL; if ((v1 >> T::SHIFT) == 2) f1 = true; else --t1; }
            }
            T::st(lookback + (size_t)tile * SIGK_BINS + d0, T::PRE | (e0 + (LB)cnt0));
            T::st(lookback + (size_t)tile * SIGK_BINS + d1, T::PRE | (e1 + (LB)cnt1));
        }
        sm.goff[d0] = (uint32_t)(bin_base[d0] + (uint64_t)e0) - base0;
        sm.goff[d1] = (uint32_t)(bin_base[d1] + (uint64_t)e1) - base1;
    };
    auto look_back = [&]() {
        if (FIRST && tid == 0 && OS_DPT == 1) {
            look_back_pair(0u, my_count[0], dbase, SIGK_SIDE_BIN, side_count, total_main);
        } else {
            if (owner) {
                uint32_t base = dbase;
#pragma unroll
                for (int k = 0; k < OS_DPT; ++k) { look_back_digit(tid * OS_DPT + k, my_count[k], base); base += my_count[k]; }
            }
            if (FIRST && tid == 0) look_back_digit(SIGK_SIDE_BIN, side_count, total_main);
        }
    };
    // The ticket of the CTA's next tile, taken as late as it can be for its keys to be requested before the write-out:
    // a tile's aggregate is published a ranking after its ticket, and every later tile's look-back waits for it.
    if (next_ticket && tid == 0) sm.next_tile = atomicAdd(next_ticket, 1u);
    if (!SIGK_OS_LATE_LOOKBACK) look_back();
    __syncthreads();

    // ---- reorder the tile in shared memory (keys, then values by the same slots)
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const uint32_t d = digit_of<FIRST>(key[i], bit_lo, digit_mask);
        const uint32_t slot = (uint32_t)wcnt16[d] + SIGK_RANK_GET(i);
        SIGK_RANK_SET(i, slot);
    }
    const uint32_t wbase = warp * (ITEMS * 32);
    // (every key of the tile has been in registers since before the first barrier: the key staging is free)
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) sm.keys[SIGK_RANK_GET(i)] = key[i];
    if (VALS_STAGED) {
        // the staged values share sm.vals with the sorted ones: everybody reads before anybody writes
        uint32_t v[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t idx = wbase + i * 32 + lane;
            v[i] = idx < tile_n ? val_at(idx) : 0u;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) sm.vals[SIGK_RANK_GET(i)] = v[i];
    } else {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t idx = wbase + i * 32 + lane;
            const uint32_t v = idx < tile_n ? val_at(idx) : 0u;
            sm.vals[SIGK_RANK_GET(i)] = v;
        }
    }

    if (SIGK_OS_LATE_LOOKBACK) look_back();
    __syncthreads();
    before_write_out();

    // ---- coalesced write-out: consecutive slots of one digit are consecutive in HBM
    for (uint32_t j = tid; j < tile_n; j += OS_THREADS) {
        const uint64_t k = sm.keys[j];
        const uint32_t d = digit_of<FIRST>(k, bit_lo, digit_mask);
        const uint32_t pos = sm.goff[d] + j;
        keys_out[pos] = k;
        vals_out[pos] = sm.vals[j];
    }
}

// L2 prefetch of a tile that a CTA of the next wave will take: cp.async.bulk.prefetch (the TMA unit pulls the lines
// into L2; nothing lands in shared memory, so it costs no staging space)
SIGK_D void prefetch_l2(const void *p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

#ifndef SIGK_OS_PREFETCH
#define SIGK_OS_PREFETCH 1
#endif

template <typename LB, bool FIRST>
__global__ void __launch_bounds__(OS_THREADS, OS_MIN_BLOCKS)
onesweep_pass_kernel(const __grid_constant__ SortSegments seg, const uint64_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                     uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, const uint64_t *__restrict__ n_ptr,
                     const uint64_t *__restrict__ off_ptr, int bit_lo, uint32_t digit_mask, const uint64_t *__restrict__ bin_base,
                     LB *__restrict__ lookback, uint32_t *__restrict__ ticket) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PassSmem &sm = *reinterpret_cast<PassSmem *>(smem_raw);
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint64_t n = FIRST ? (n_ptr ? *n_ptr : seg.start[seg.n]) : *n_ptr;
    const uint64_t off = (!FIRST && off_ptr) ? *off_ptr : 0ull;      // the run's place in the in / out buffers
    const uint32_t wbase = warp * (OS_ITEMS * 32);
    if (!FIRST && SIGK_OS_PIPELINE) {
        // ---- the plain pass, software-pipelined over the tiles of a persistent CTA: the ticket of the next tile is taken
        // while this one is ranked, and its keys are requested before this tile's write-out (the key registers are dead
        // by then), so their latency hides behind the write-out instead of standing in front of the ranking
        uint64_t key[OS_ITEMS];
        auto load_keys = [&](uint32_t t) {
            const uint64_t t0 = (uint64_t)t * OS_TILE;
            const uint32_t tn = (uint32_t)((n - t0) < (uint64_t)OS_TILE ? (n - t0) : (uint64_t)OS_TILE);
            // Padding of the last tile gets key ~0: it ranks after every real record of the top digit and is never written.
#pragma unroll
            for (int i = 0; i < OS_ITEMS; ++i) {
                const uint32_t idx = wbase + i * 32 + lane;
                key[i] = idx < tn ? ld_stream_u64(keys_in + off + t0 + idx) : PAD_KEY;
            }
        };
        if (tid == 0) sm.tile = atomicAdd(ticket, 1u);
        os_zero_counters(sm);
        __syncthreads();
        uint32_t tile = sm.tile;
        if ((uint64_t)tile * OS_TILE >= n) return;
        load_keys(tile);
        for (;;) {
            const uint64_t tile_start = (uint64_t)tile * OS_TILE;
            const uint32_t tile_n = (uint32_t)((n - tile_start) < (uint64_t)OS_TILE ? (n - tile_start) : (uint64_t)OS_TILE);
            if (tid == 0) {
                if (!SIGK_OS_PERSISTENT) sm.next_tile = 0xFFFFFFFFu;
#if SIGK_OS_PREFETCH
                // L2 prefetch of the tile one wave ahead (gridDim.x CTAs are resident)
                const uint64_t ahead = tile_start + (uint64_t)gridDim.x * OS_TILE;
                if (ahead + OS_TILE <= n && (((uintptr_t)(keys_in + off + ahead) | (uintptr_t)(vals_in + off + ahead)) & 15u) == 0) {
                    prefetch_l2(keys_in + off + ahead, OS_TILE * sizeof(uint64_t));
                    prefetch_l2(vals_in + off + ahead, OS_TILE * sizeof(uint32_t));
                }
#endif
            }
            uint32_t next = 0xFFFFFFFFu;
            os_sort_tile<LB, false, false, OS_ITEMS>(sm, key, tile, tile_n, bit_lo, digit_mask, bin_base, lookback,
                                                     [&](uint32_t r) { return ld_stream_u32(vals_in + off + tile_start + r); },
                                                     keys_out + off, vals_out + off, SIGK_OS_PERSISTENT ? ticket : nullptr, [&]() {
                                                         next = sm.next_tile;
                                                         os_zero_counters(sm);
                                                         if ((uint64_t)next * OS_TILE < n) load_keys(next);
                                                     });
            __syncthreads();                                // the write-out is done with the staging; the counters are zero
            if ((uint64_t)next * OS_TILE >= n) return;
            tile = next;
        }
    }
    for (;;) {
        __syncthreads();                                    // the previous tile's write-out is done with the staging
        if (tid == 0) sm.tile = atomicAdd(ticket, 1u);
        os_zero_counters(sm);
        __syncthreads();
        const uint32_t tile = sm.tile;
        uint64_t key[OS_ITEMS];
        if (!FIRST) {
            const uint64_t tile_start = (uint64_t)tile * OS_TILE;
            if (tile_start >= n) return;
            const uint32_t tile_n = (uint32_t)((n - tile_start) < (uint64_t)OS_TILE ? (n - tile_start) : (uint64_t)OS_TILE);
#if SIGK_OS_PREFETCH
            if (tid == 0) {
                // L2 prefetch of the tile one wave ahead (gridDim.x CTAs are resident)
                const uint64_t ahead = tile_start + (uint64_t)gridDim.x * OS_TILE;
                if (ahead + OS_TILE <= n && (((uintptr_t)(keys_in + off + ahead) | (uintptr_t)(vals_in + off + ahead)) & 15u) == 0) {
                    prefetch_l2(keys_in + off + ahead, OS_TILE * sizeof(uint64_t));
                    prefetch_l2(vals_in + off + ahead, OS_TILE * sizeof(uint32_t));
                }
            }
#endif
            // Padding of the last tile gets key ~0: it ranks after every real record of the top digit and is never written.
#pragma unroll
            for (int i = 0; i < OS_ITEMS; ++i) {
                const uint32_t idx = wbase + i * 32 + lane;
                key[i] = idx < tile_n ? ld_stream_u64(keys_in + off + tile_start + idx) : PAD_KEY;
            }
            os_sort_tile<LB, false, false, OS_ITEMS>(sm, key, tile, tile_n, bit_lo, digit_mask, bin_base, lookback,
                                                     [&](uint32_t r) { return ld_stream_u32(vals_in + off + tile_start + r); },
                                                     keys_out + off, vals_out + off, nullptr, []() {});
        } else {
            // The regions of the first pass are read in place, one after the other, in tiles that never straddle two of
            // them (every region ends with a short tile): a tile is one pointer pair and a count, as in the plain pass.
            if (tile >= (seg.n == 1 && n_ptr ? (uint32_t)((n + OS_TILE - 1) / OS_TILE) : seg.tile_start[seg.n])) return;
            int r = 0;
            while (r + 1 < seg.n && tile >= seg.tile_start[r + 1]) ++r;
            const uint64_t rec0 = (uint64_t)(tile - seg.tile_start[r]) * OS_TILE;
            const uint64_t n_r = (seg.n == 1 && n_ptr) ? n : seg.start[r + 1] - seg.start[r];
            const uint32_t tile_n = (uint32_t)((n_r - rec0) < (uint64_t)OS_TILE ? (n_r - rec0) : (uint64_t)OS_TILE);
            const uint64_t *kp = seg.keys[r] + rec0;
            const uint32_t *vp = seg.vals[r] + rec0;
#pragma unroll
            for (int i = 0; i < OS_ITEMS; ++i) {
                const uint32_t idx = wbase + i * 32 + lane;
                key[i] = idx < tile_n ? ld_stream_u64(kp + idx) : PAD_KEY;
            }
            os_sort_tile<LB, true, false, OS_ITEMS>(sm, key, tile, tile_n, bit_lo, digit_mask, bin_base, lookback,
                                                    [&](uint32_t idx) { return ld_stream_u32(vp + idx); }, keys_out, vals_out, nullptr, []() {});
        }
    }
}

// ---- encode fused with the first pass -------------------------------------------------------------------
// A tile is ES_WARPS slices of 512 window positions.  Warps 0..ES_WARPS-1 run the window loop (window_scan.cuh)
// and compact the valid windows of the tile into the staging in canonical order; then all warps read the staging
// warp-striped and the tile goes through os_sort_tile like any other: main records to their lowest digit's run,
// records with a lower-case residue to the side bin.
// The window loop of one fused tile: the valid windows of the tile's ES_WARPS slices, compacted into the staging in
// canonical order (one pad slot per 16: lanes write runs of ~16 records).  Returns the number of records.  Inlined:
// the kernel then spills ~200 bytes per thread (to L2: 220 KB of the SM's 228 KB are shared memory), and is still
// faster than with this out of line, where each half has the whole register budget but the call saves and restores
// around it (config 2, measured back to back: 8.16 ms inlined, 8.50 ms out of line; SIGK_ES_NOINLINE=1 builds the latter).
#ifndef SIGK_ES_NOINLINE
#define SIGK_ES_NOINLINE 0
#endif
#if SIGK_ES_NOINLINE
#define SIGK_ES_INLINE __noinline__
#else
#define SIGK_ES_INLINE __forceinline__
#endif
template <typename SM>
__device__ SIGK_ES_INLINE uint32_t encode_tile_to_staging(SM &sm, const EncodeArgs &a, uint32_t tile) {
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t sub = tile * ES_WARPS + warp;
    const bool enc = warp < ES_WARPS && (uint64_t)sub * WS_SUB < a.total_res;
    WindowLane w;
    uint32_t valid = 0, incl = 0, mine = 0;
    if (enc) {
        ws_load(a, sm.sym, sub, w);
        valid = ws_valid_mask(a, w);
        mine = incl = (uint32_t)__popc(valid);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (unsigned)o) incl += y;
        }
    }
    if (lane == 31) sm.wtotal[warp] = incl;             // 0 for warps without a slice
    __syncthreads();
    uint32_t before = 0, tile_n = 0;
#pragma unroll
    for (int q = 0; q < ES_WARPS; ++q) {
        const uint32_t c = sm.wtotal[q];
        before += q < (int)warp ? c : 0u;
        tile_n += c;
    }
    if (enc) {
        uint32_t o = before + incl - mine;
        ws_for_each(a, w, valid, [&](int, uint64_t key, uint32_t i) {
            const uint32_t slot = stage_slot(o++);
            sm.keys[slot] = key;
            sm.vals[slot] = a.ordinal_base + i;
        });
    }
    __syncthreads();
    return tile_n;
}

template <typename LB>
__global__ void __launch_bounds__(OS_THREADS, OS_MIN_BLOCKS)
encode_sort_kernel(EncodeArgs a, uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int bit_lo, uint32_t digit_mask,
                   const uint64_t *__restrict__ bin_base, LB *__restrict__ lookback, uint32_t *__restrict__ ticket, uint32_t n_tiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FusedSmem &sm = *reinterpret_cast<FusedSmem *>(smem_raw);
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    ws_fill_symbols(sm.sym);
    const uint32_t wbase = warp * (ES_ITEMS * 32);
    for (;;) {
        __syncthreads();                                    // the previous tile's write-out is done with the staging
        if (tid == 0) sm.tile = atomicAdd(ticket, 1u);
        os_zero_counters(sm);
        __syncthreads();
        const uint32_t tile = sm.tile;
        if (tile >= n_tiles) return;

        const uint32_t tile_n = encode_tile_to_staging(sm, a, tile);
        uint64_t key[ES_ITEMS];
#pragma unroll
        for (int i = 0; i < ES_ITEMS; ++i) {
            const uint32_t idx = wbase + i * 32 + lane;
            key[i] = idx < tile_n ? sm.keys[stage_slot(idx)] : PAD_KEY;
        }
        os_sort_tile<LB, true, true, ES_ITEMS>(sm, key, tile, tile_n, bit_lo, digit_mask, bin_base, lookback,
                                               [&](uint32_t idx) { return sm.vals[stage_slot(idx)]; }, keys_out, vals_out, nullptr, []() {});
    }
}

// ---- encode and route (multi-GPU) ----------------------------------------------------------------------------
// The same tile machinery with the OWNER of a record's k-mer range as the digit: the valid windows of 14 slices are
// compacted in canonical order, ranked by owner (the owner rides in the five spare low bits of the staged key, so it
// is computed once), reordered in shared memory and written as one contiguous run per owner — ~900 records, 7 KB of
// keys — straight into region `this rank` of the owner's landing zone (peer memory over NVLink), or into a local
// send region.  A run's place in its region is a chained scan per owner over the tiles, in ticket = canonical order,
// so every region holds its records in insertion order.  (The round-1 kernel routed per 512-position warp slice:
// 64-record runs, eight times the scan entries, ~485 GB/s of NVLink egress.)
constexpr uint32_t RT_BINS = 32;                    // owners 0..15, 31 = rows of the tile that hold no record

__global__ void __launch_bounds__(OS_THREADS, OS_MIN_BLOCKS)
encode_route_kernel(EncodeArgs a, const __grid_constant__ EncodeSplitArgs sp, uint32_t *__restrict__ ticket, uint32_t n_tiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FusedSmem &sm = *reinterpret_cast<FusedSmem *>(smem_raw);
    __shared__ uint64_t s_split[16];
    __shared__ uint64_t *s_dkeys[16];
    __shared__ uint32_t *s_dvals[16];
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const uint32_t W = (uint32_t)sp.n_split + 1u;
    ws_fill_symbols(sm.sym);
    if (tid < (unsigned)sp.n_split) s_split[tid] = sp.split_codes[tid];
    if (tid < 16) { s_dkeys[tid] = sp.dst_keys[tid]; s_dvals[tid] = sp.dst_vals[tid]; }
    const uint32_t wbase = warp * (ES_ITEMS * 32);
    uint16_t *wcnt16 = reinterpret_cast<uint16_t *>(sm.cnt[warp]);
    for (;;) {
        __syncthreads();                                    // the previous tile's write-out is done with the staging
        if (tid == 0) sm.tile = atomicAdd(ticket, 1u);
        if (tid < OS_WARPS * (RT_BINS / 2)) sm.cnt[tid / (RT_BINS / 2)][tid % (RT_BINS / 2)] = 0;
        __syncthreads();
        const uint32_t tile = sm.tile;
        if (tile >= n_tiles) return;
        const bool last_tile = tile + 1 == n_tiles;

        // ---- window loop, part 1: which windows are valid, how many per slice
        const uint32_t sub = tile * ES_WARPS + warp;
        const bool enc = warp < ES_WARPS && (uint64_t)sub * WS_SUB < a.total_res;
        WindowLane w;
        uint32_t valid = 0, incl = 0, mine = 0;
        if (enc) {
            ws_load(a, sm.sym, sub, w);
            valid = ws_valid_mask(a, w);
            mine = incl = (uint32_t)__popc(valid);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= (unsigned)o) incl += y;
            }
        }
        if (lane == 31) sm.wtotal[warp] = incl;             // 0 for warps without a slice
        __syncthreads();
        uint32_t before = 0, tile_n = 0;
#pragma unroll
        for (int q = 0; q < ES_WARPS; ++q) {
            const uint32_t c = sm.wtotal[q];
            before += q < (int)warp ? c : 0u;
            tile_n += c;
        }
        // ---- part 2: the records, compacted in canonical order; the owner (number of splitter codes <= the record's
        // case-folded code) goes into the key's spare low bits
        if (enc) {
            uint32_t o = before + incl - mine;
            ws_for_each_code(a, w, valid, [&](int, uint64_t code, uint32_t mask, uint32_t off, uint32_t i) {
                uint32_t d = 0;
                for (int k = 0; k < sp.n_split; ++k) d += code >= s_split[k] ? 1u : 0u;
                const uint32_t slot = stage_slot(o++);
                sm.keys[slot] = sigk_pack_key(code, mask, off) | d;
                sm.vals[slot] = a.ordinal_base + i;
            }, [&](uint32_t i, uint32_t n_windows) { if (a.prot_windows) atomicAdd(a.prot_windows + i, n_windows); });
        }
        __syncthreads();
        uint64_t key[ES_ITEMS];
#pragma unroll
        for (int i = 0; i < ES_ITEMS; ++i) {
            const uint32_t idx = wbase + i * 32 + lane;
            key[i] = idx < tile_n ? sm.keys[stage_slot(idx)] : PAD_KEY;         // the low bits of PAD_KEY: bin 31, behind every owner
        }
        // ---- rank by owner inside the warp (five ballots), as in os_sort_tile
        uint32_t rank2[(ES_ITEMS + 1) / 2];
#pragma unroll
        for (int i = 0; i < (ES_ITEMS + 1) / 2; ++i) rank2[i] = 0;
#pragma unroll
        for (int i = 0; i < ES_ITEMS; ++i) {
            const uint32_t d = (uint32_t)key[i] & (RT_BINS - 1);
            unsigned peers = 0xffffffffu;
#pragma unroll
            for (int b = 0; b < 5; ++b) peers_step(peers, d, 1u << b);
            const uint32_t below = (uint32_t)__popc(peers & lt);
            const uint32_t old = wcnt16[d];
            if (below == 0) wcnt16[d] = (uint16_t)(old + (uint32_t)__popc(peers));
            SIGK_RANK_SET(i, old + below);
            __syncwarp();
        }
        __syncthreads();
        // ---- warp 0, lane d = owner d: tile count, chained scan over the tiles, the run's place in the owner's region
        if (warp == 0) {
            uint32_t cnt = 0;
#pragma unroll
            for (int q = 0; q < OS_WARPS; ++q) cnt += reinterpret_cast<const uint16_t *>(sm.cnt[q])[lane];
            uint64_t *state = sp.owner_state + (size_t)tile * RT_BINS + lane;
            if (lane < W) st_volatile_u64(state, (tile == 0 ? SIGK_CS_PRE : SIGK_CS_AGG) | (uint64_t)cnt);
            uint32_t incl_d = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, incl_d, o);
                if (lane >= (unsigned)o) incl_d += y;
            }
            const uint32_t dbase = incl_d - cnt;            // first sorted slot of owner `lane` in the tile
            uint32_t run = dbase;
#pragma unroll
            for (int q = 0; q < OS_WARPS; ++q) {
                uint16_t *c = reinterpret_cast<uint16_t *>(sm.cnt[q]) + lane;
                const uint32_t x = *c;
                *c = (uint16_t)run;
                run += x;
            }
            uint64_t excl = 0;
            if (lane < W) {
                if (tile > 0) {
                    int64_t t = (int64_t)tile - 1;
                    for (;;) {
                        const uint64_t v = ld_volatile_u64(sp.owner_state + (size_t)t * RT_BINS + lane);
                        const uint64_t flag = v >> 62;
                        if (flag == 0) continue;
                        excl += v & SIGK_CS_VAL;
                        if (flag == 2) break;
                        --t;
                    }
                    st_volatile_u64(state, SIGK_CS_PRE | (excl + cnt));
                }
                if (excl + cnt > sp.region_stride) atomicOr(sp.overflow, 1u);
                if (last_tile) sp.owner_totals[lane] = excl + cnt;
            }
            sm.goff[lane] = (uint32_t)excl - dbase;         // region offsets fit 32 bits (a region holds < 2^32 records)
        }
        __syncthreads();
        // ---- reorder in shared memory: keys, then the staged values (everybody reads them before anybody writes)
#pragma unroll
        for (int i = 0; i < ES_ITEMS; ++i) {
            const uint32_t slot = (uint32_t)wcnt16[(uint32_t)key[i] & (RT_BINS - 1)] + SIGK_RANK_GET(i);
            SIGK_RANK_SET(i, slot);
        }
#pragma unroll
        for (int i = 0; i < ES_ITEMS; ++i) sm.keys[SIGK_RANK_GET(i)] = key[i];
        uint32_t v[ES_ITEMS];
#pragma unroll
        for (int i = 0; i < ES_ITEMS; ++i) {
            const uint32_t idx = wbase + i * 32 + lane;
            v[i] = idx < tile_n ? sm.vals[stage_slot(idx)] : 0u;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ES_ITEMS; ++i) sm.vals[SIGK_RANK_GET(i)] = v[i];
        __syncthreads();
        if (*reinterpret_cast<volatile uint32_t *>(sp.overflow)) continue;     // regions too small: the caller grows them and encodes again
        // ---- write-out: one contiguous run per owner, coalesced (a warp's store is 256 bytes of keys)
        for (uint32_t j = tid; j < tile_n; j += OS_THREADS) {
            const uint64_t k = sm.keys[j];
            const uint32_t d = (uint32_t)k & (RT_BINS - 1);
            const uint32_t pos = sm.goff[d] + j;
            s_dkeys[d][pos] = k & ~(uint64_t)(RT_BINS - 1);
            s_dvals[d][pos] = sm.vals[j];
        }
    }
}

// ---- histograms of already encoded records ------------------------------------------------------------------
constexpr int HIST_THREADS = 512;
constexpr int HIST_UNROLL = 4;

// MAIN: the input of a first pass (regions read in place); records with mask8 != 0 only bump the side bin of row 0.
template <bool MAIN>
__global__ void __launch_bounds__(HIST_THREADS)
histogram_kernel(const __grid_constant__ SortSegments seg, const uint64_t *__restrict__ keys, const uint64_t *__restrict__ n_ptr,
                 const uint64_t *__restrict__ off_ptr, PassPlan plan, uint64_t *__restrict__ hist) {
    __shared__ uint32_t sh[SORT_MAX_PASSES * SIGK_RADIX];
    __shared__ uint32_t s_side;
    for (int j = threadIdx.x; j < plan.npass * SIGK_RADIX; j += HIST_THREADS) sh[j] = 0;
    if (threadIdx.x == 0) s_side = 0;
    __syncthreads();
    const uint64_t n = MAIN ? (n_ptr ? *n_ptr : seg.start[seg.n]) : *n_ptr;
    if (!MAIN && off_ptr) keys += *off_ptr;
    uint32_t side = 0;
    const uint64_t stride = (uint64_t)gridDim.x * HIST_THREADS * HIST_UNROLL;
    for (uint64_t base = (uint64_t)blockIdx.x * HIST_THREADS * HIST_UNROLL; base < n; base += stride) {
        uint64_t k[HIST_UNROLL];
        bool ok[HIST_UNROLL];
        int r_base = 0;
        if (MAIN) while (r_base + 1 < seg.n && base >= seg.start[r_base + 1]) ++r_base;
#pragma unroll
        for (int u = 0; u < HIST_UNROLL; ++u) {
            const uint64_t idx = base + (uint64_t)u * HIST_THREADS + threadIdx.x;
            ok[u] = idx < n;
            k[u] = 0;
            if (ok[u]) {
                if (MAIN) {
                    // the region of the chunk's first record is found once per chunk; records past its end walk on
                    int r = r_base;
                    while (r + 1 < seg.n && idx >= seg.start[r + 1]) ++r;
                    k[u] = ld_stream_u64(seg.keys[r] + (idx - seg.start[r]));
                } else {
                    k[u] = ld_stream_u64(keys + idx);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < HIST_UNROLL; ++u) {
            if (!ok[u]) continue;
            if (MAIN && sigk_key_mask(k[u])) { ++side; continue; }
            for (int p = 0; p < plan.npass; ++p)
                atomicAdd(&sh[p * SIGK_RADIX + ((uint32_t)(k[u] >> plan.lo[p]) & ((1u << plan.bits[p]) - 1u))], 1u);
        }
    }
    if (MAIN) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) side += __shfl_xor_sync(0xffffffffu, side, o);
        if ((threadIdx.x & 31u) == 0 && side) atomicAdd(&s_side, side);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < plan.npass * SIGK_RADIX; j += HIST_THREADS)
        if (sh[j]) atomicAdd(reinterpret_cast<unsigned long long *>(hist + (size_t)(j / SIGK_RADIX) * SIGK_BINS + (j % SIGK_RADIX)),
                             (unsigned long long)sh[j]);
    if (MAIN && threadIdx.x == 0 && s_side) atomicAdd(reinterpret_cast<unsigned long long *>(hist + SIGK_SIDE_BIN), (unsigned long long)s_side);
}

// bin_base[p][d] = records of pass p with a smaller digit; row 0 continues into the side bin
__global__ void scan_bins_kernel(const uint64_t *__restrict__ hist, uint64_t *__restrict__ bin_base, uint64_t *__restrict__ counts) {
    __shared__ uint64_t s[SIGK_RADIX];
    const int p = blockIdx.x, d = threadIdx.x;
    s[d] = hist[(size_t)p * SIGK_BINS + d];
    __syncthreads();
    // 512 bins: a serial scan by one thread is a few hundred cycles
    if (d == 0) {
        uint64_t run = 0;
        for (int i = 0; i < SIGK_RADIX; ++i) { const uint64_t c = s[i]; s[i] = run; run += c; }
        const uint64_t side = hist[(size_t)p * SIGK_BINS + SIGK_SIDE_BIN];
        bin_base[(size_t)p * SIGK_BINS + SIGK_SIDE_BIN] = run;
        if (p == 0 && counts) { counts[0] = run; counts[1] = side; counts[2] = run + side; }
    }
    __syncthreads();
    bin_base[(size_t)p * SIGK_BINS + d] = s[d];
}

template <typename LB>
__global__ void clear_lookback_kernel(LB *__restrict__ lookback, const uint64_t *__restrict__ n_ptr, uint32_t tile_records) {
    const uint64_t words = ((*n_ptr + tile_records - 1) / tile_records) * SIGK_BINS;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (uint64_t)gridDim.x * blockDim.x) lookback[i] = 0;
}

}  // namespace

static bool wide_lookback(uint64_t capacity) { return capacity >= (1ull << 30); }

size_t onesweep_lookback_bytes(uint64_t capacity) {
    // rows for the smaller (fused) tile cover both kinds of pass
    // (+ one short tile per region of a first pass over several regions)
    return (size_t)(encode_sort_tiles(capacity) + 1 + SORT_MAX_SEGMENTS) * SIGK_BINS * (wide_lookback(capacity) ? 8 : 4);
}

cudaError_t onesweep_configure() {
    cudaError_t e;
#define SIGK_OPT_IN(kernel, bytes) \
    if ((e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))) != cudaSuccess) return e;
    SIGK_OPT_IN((onesweep_pass_kernel<uint32_t, false>), sizeof(PassSmem))
    SIGK_OPT_IN((onesweep_pass_kernel<uint64_t, false>), sizeof(PassSmem))
    SIGK_OPT_IN((onesweep_pass_kernel<uint32_t, true>), sizeof(PassSmem))
    SIGK_OPT_IN((onesweep_pass_kernel<uint64_t, true>), sizeof(PassSmem))
    SIGK_OPT_IN(encode_sort_kernel<uint32_t>, sizeof(FusedSmem))
    SIGK_OPT_IN(encode_sort_kernel<uint64_t>, sizeof(FusedSmem))
    SIGK_OPT_IN(encode_route_kernel, sizeof(FusedSmem))
#undef SIGK_OPT_IN
    return cudaSuccess;
}

static unsigned persistent_grid(uint64_t tiles, int sm_count) {
    if (!SIGK_OS_PERSISTENT) return (unsigned)std::max<uint64_t>(1, tiles);       // every CTA finds its ticket in range once, then leaves
    return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(tiles, (uint64_t)sm_count * OS_MIN_BLOCKS));
}

cudaError_t launch_histogram(const uint64_t *keys, const uint64_t *n_ptr, const uint64_t *off_ptr, uint64_t capacity,
                             const PassPlan &plan, uint64_t *hist, int sm_count, cudaStream_t stream) {
    if (capacity == 0) return cudaSuccess;
    uint64_t want = (capacity + (uint64_t)HIST_THREADS * HIST_UNROLL - 1) / ((uint64_t)HIST_THREADS * HIST_UNROLL);
    const uint64_t cap = (uint64_t)sm_count * 4;
    if (want > cap) want = cap;
    histogram_kernel<false><<<(unsigned)want, HIST_THREADS, 0, stream>>>(SortSegments{}, keys, n_ptr, off_ptr, plan, hist);
    return cudaGetLastError();
}

cudaError_t launch_histogram_main(const SortSegments &seg, const uint64_t *n_ptr, const PassPlan &plan, uint64_t *hist,
                                  int sm_count, cudaStream_t stream) {
    if (seg.n < 1 || seg.n > SORT_MAX_SEGMENTS) return cudaErrorInvalidValue;
    histogram_kernel<true><<<(unsigned)sm_count * 4, HIST_THREADS, 0, stream>>>(seg, nullptr, n_ptr, nullptr, plan, hist);
    return cudaGetLastError();
}

cudaError_t launch_scan_bins(const uint64_t *hist, uint64_t *bin_base, int npass, uint64_t *counts, cudaStream_t stream) {
    scan_bins_kernel<<<npass, SIGK_RADIX, 0, stream>>>(hist, bin_base, counts);
    return cudaGetLastError();
}

cudaError_t launch_clear_lookback(void *lookback, const uint64_t *n_ptr, uint64_t capacity, cudaStream_t stream) {
    if (capacity == 0) return cudaSuccess;
    const uint64_t words = onesweep_tiles(capacity) * SIGK_BINS;
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((words + 255) / 256, 1184));
    if (wide_lookback(capacity)) clear_lookback_kernel<uint64_t><<<grid, 256, 0, stream>>>((uint64_t *)lookback, n_ptr, (uint32_t)OS_TILE);
    else clear_lookback_kernel<uint32_t><<<grid, 256, 0, stream>>>((uint32_t *)lookback, n_ptr, (uint32_t)OS_TILE);
    return cudaGetLastError();
}

cudaError_t launch_onesweep_pass(const uint64_t *keys_in, const uint32_t *vals_in, uint64_t *keys_out,
                                 uint32_t *vals_out, const uint64_t *n_ptr, const uint64_t *off_ptr, uint64_t capacity, int bit_lo, int nbits,
                                 const uint64_t *bin_base, void *lookback, uint32_t *ticket, int sm_count, cudaStream_t stream) {
    if (capacity == 0) return cudaSuccess;
    const unsigned grid = persistent_grid(onesweep_tiles(capacity), sm_count);
    const uint32_t mask = (1u << nbits) - 1u;
    if (wide_lookback(capacity))
        onesweep_pass_kernel<uint64_t, false><<<grid, OS_THREADS, sizeof(PassSmem), stream>>>(
            SortSegments{}, keys_in, vals_in, keys_out, vals_out, n_ptr, off_ptr, bit_lo, mask, bin_base, (uint64_t *)lookback, ticket);
    else
        onesweep_pass_kernel<uint32_t, false><<<grid, OS_THREADS, sizeof(PassSmem), stream>>>(
            SortSegments{}, keys_in, vals_in, keys_out, vals_out, n_ptr, off_ptr, bit_lo, mask, bin_base, (uint32_t *)lookback, ticket);
    return cudaGetLastError();
}

cudaError_t launch_onesweep_first_pass(const SortSegments &seg, const uint64_t *n_ptr, uint64_t *keys_out, uint32_t *vals_out,
                                       uint64_t capacity, int bit_lo, int nbits, const uint64_t *bin_base, void *lookback,
                                       uint32_t *ticket, int sm_count, cudaStream_t stream) {
    if (capacity == 0) return cudaSuccess;
    if (seg.n < 1 || seg.n > SORT_MAX_SEGMENTS) return cudaErrorInvalidValue;
    SortSegments sg = seg;
    sg.tile_start[0] = 0;                                  // region-aligned tiles: region r owns tiles tile_start[r] .. tile_start[r+1]
    for (int r = 0; r < sg.n; ++r) sg.tile_start[r + 1] = sg.tile_start[r] + (uint32_t)onesweep_tiles(sg.start[r + 1] - sg.start[r]);
    // (the look-back rows are sized for capacity / ES_TILE + 1 tiles: up to one short tile per region more than a plain pass)
    const unsigned grid = persistent_grid(onesweep_tiles(capacity) + SORT_MAX_SEGMENTS, sm_count);
    const uint32_t mask = (1u << nbits) - 1u;
    if (wide_lookback(capacity))
        onesweep_pass_kernel<uint64_t, true><<<grid, OS_THREADS, sizeof(PassSmem), stream>>>(
            sg, nullptr, nullptr, keys_out, vals_out, n_ptr, nullptr, bit_lo, mask, bin_base, (uint64_t *)lookback, ticket);
    else
        onesweep_pass_kernel<uint32_t, true><<<grid, OS_THREADS, sizeof(PassSmem), stream>>>(
            sg, nullptr, nullptr, keys_out, vals_out, n_ptr, nullptr, bit_lo, mask, bin_base, (uint32_t *)lookback, ticket);
    return cudaGetLastError();
}

cudaError_t launch_encode_sort(const EncodeArgs &a, uint64_t *keys_out, uint32_t *vals_out, int bit_lo, int nbits,
                               const uint64_t *bin_base, void *lookback, uint32_t *ticket, int sm_count, cudaStream_t stream) {
    if (a.total_res == 0 || a.n_prot == 0) return cudaSuccess;
    const uint64_t tiles = encode_sort_tiles(a.total_res);
    const unsigned grid = persistent_grid(tiles, sm_count);
    const uint32_t mask = (1u << nbits) - 1u;
    if (wide_lookback(a.total_res))
        encode_sort_kernel<uint64_t><<<grid, OS_THREADS, sizeof(FusedSmem), stream>>>(a, keys_out, vals_out, bit_lo, mask, bin_base,
                                                                                        (uint64_t *)lookback, ticket, (uint32_t)tiles);
    else
        encode_sort_kernel<uint32_t><<<grid, OS_THREADS, sizeof(FusedSmem), stream>>>(a, keys_out, vals_out, bit_lo, mask, bin_base,
                                                                                        (uint32_t *)lookback, ticket, (uint32_t)tiles);
    return cudaGetLastError();
}

cudaError_t launch_encode_route(const EncodeArgs &a, const EncodeSplitArgs &sp, uint32_t *ticket, int sm_count, cudaStream_t stream) {
    if (a.total_res == 0 || a.n_prot == 0) return cudaSuccess;
    if (sp.n_split < 0 || sp.n_split > 15) return cudaErrorInvalidValue;
    const uint64_t tiles = encode_sort_tiles(a.total_res);
    encode_route_kernel<<<persistent_grid(tiles, sm_count), OS_THREADS, sizeof(FusedSmem), stream>>>(a, sp, ticket, (uint32_t)tiles);
    return cudaGetLastError();
}

}  // namespace sigk
