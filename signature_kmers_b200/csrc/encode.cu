// encode.cu — stage 1: slide the 8-residue window over the packed protein
// buffer, drop windows the reference drops, emit one 12-byte record per valid
// window in canonical (insertion) order.
//
// Replaces the window loop of SignatureBuilder<K>::load_kmers_from_sequence
// (reference src/signature_build.tcc:162-180):
//   - a window is valid iff all 8 bytes are in ok_prot_ (src/signature_build.h:102-103)
//     and it lies inside one protein (it < seq.end()-K+1);
//   - offset = (unsigned short)(len - p)   (:164);
//   - records are emitted in increasing position, proteins in input order, which
//     is the multimap insertion order of the reference's serial branch (:50-56).
//
// Layout: each thread owns 16 consecutive window positions and reads its 16
// residues with one 128-bit load (a warp reads 512 contiguous bytes); the 7
// look-ahead residues come from the next lane by shuffle.  The 43-bit base-40
// code is rolled (one multiply-add per window).  Valid windows are compacted
// through shared memory and leave the CTA as coalesced 8-byte / 4-byte stores;
// the CTA's place in the output is a single-pass chained scan over tile totals
// (tiles are tickets, so the record order is the position order); every warp
// owns its own slice and scan entry, so there is no block barrier after the ticket.
//
// HBM traffic per valid window: ~1 byte read, 12 bytes written.
#include "kernels.h"
#include "sigk_common.cuh"

#include <cstdlib>

namespace sigk {

namespace {

constexpr int ENC_WARPS = ENC_THREADS / 32;
constexpr int ENC_SUB = 32 * ENC_PPT;                 // window positions one warp owns: 512
// compacted records are staged with one pad slot per 16 so that the
// thread-contiguous writes (stride ~16 records between lanes) spread over banks
constexpr int ENC_STAGE = ENC_SUB + ENC_SUB / 16;
SIGK_D int stage_slot(int o) { return o + (o >> 4); }

struct EncSmem {
    uint64_t keys[ENC_WARPS][ENC_STAGE];
    uint32_t vals[ENC_WARPS][ENC_STAGE];
    uint32_t tile;
};

// largest i in [lo, hi] with starts[i] <= g   (starts[lo] <= g is guaranteed)
SIGK_D uint32_t find_protein(const uint64_t *__restrict__ starts, uint32_t lo, uint32_t hi, uint64_t g) {
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo + 1) >> 1);
        if (__ldg(starts + mid) <= g) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

// The warps of a CTA are independent after the ticket: each owns ENC_SUB positions and
// its own entry in the chained scan, so nothing waits at a block barrier.
__global__ void __launch_bounds__(ENC_THREADS, 4)
encode_kernel(EncodeArgs a, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals,
              uint64_t *__restrict__ scan_state, uint32_t *__restrict__ ticket, uint64_t *__restrict__ n_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EncSmem &sm = *reinterpret_cast<EncSmem *>(smem_raw);

    __shared__ int8_t s_sym[256];
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) sm.tile = atomicAdd(ticket, 1u);
    static_assert(ENC_THREADS == 256, "one table entry per thread");
    s_sym[tid] = (int8_t)sigk_symbol(tid);
    __syncthreads();
    const uint32_t sub = sm.tile * ENC_WARPS + warp;                 // this warp's entry in the chained scan
    const uint64_t g0 = (uint64_t)sub * ENC_SUB;
    if (g0 >= a.total_res) return;
    const bool last_sub = g0 + ENC_SUB >= a.total_res;

    // protein range of the warp's positions, from the per-slice index built by slice_index_kernel
    const uint32_t p_lo = __ldg(a.slice_prot + sub);
    const uint32_t p_hi = last_sub ? a.n_prot - 1 : __ldg(a.slice_prot + sub + 1);

    // residues: 16 of my own + 8 of the next lane's
    const uint64_t g_first = g0 + (uint64_t)lane * ENC_PPT;
    const uint4 w = ld_stream_u128(reinterpret_cast<const uint4 *>(a.res + g_first));
    uint32_t n0 = __shfl_down_sync(0xffffffffu, w.x, 1);
    uint32_t n1 = __shfl_down_sync(0xffffffffu, w.y, 1);
    if (lane == 31) {
        const uint2 nx = *reinterpret_cast<const uint2 *>(a.res + g_first + ENC_PPT);
        n0 = nx.x; n1 = nx.y;
    }
    const uint32_t words[6] = {w.x, w.y, w.z, w.w, n0, n1};

    // symbols stay packed four to a register (23 ints would not fit the register budget)
    uint32_t sw[6] = {0, 0, 0, 0, 0, 0};
    uint32_t bad = 0;
#pragma unroll
    for (int j = 0; j < ENC_PPT + 7; ++j) {
        const unsigned c = (words[j >> 2] >> (8 * (j & 3))) & 0xFFu;
        int sy = s_sym[c];                              // byte -> symbol 0..39, or -1 (table: 1/5 of the arithmetic form's instructions)
        if (sy < 0) { bad |= 1u << j; sy = 0; }
        sw[j >> 2] |= (uint32_t)sy << (8 * (j & 3));
    }
#define SIGK_SYM(j) ((uint64_t)((sw[(j) >> 2] >> (8 * ((j) & 3))) & 0xFFu))

    // pass A: which of my 16 windows are valid
    const uint32_t p_first = find_protein(a.starts, p_lo, p_hi, g_first < a.total_res ? g_first : a.total_res - 1);
    uint32_t valid_mask = 0;
    {
        uint32_t i = p_first;
        uint64_t prot_end = __ldg(a.starts + i + 1);
#pragma unroll
        for (int j = 0; j < ENC_PPT; ++j) {
            const uint64_t g = g_first + j;
            while (g >= prot_end && i + 1 < a.n_prot) { ++i; prot_end = __ldg(a.starts + i + 1); }
            if (((bad >> j) & 0xFFu) == 0 && g + SIGK_K_DEV <= prot_end) valid_mask |= 1u << j;
        }
    }
    const uint32_t mine = (uint32_t)__popc(valid_mask);
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += y;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    chained_scan_publish_warp(scan_state, sub, total);     // resolved after pass B: the predecessors publish meanwhile

    // pass B: roll the code, emit valid windows to their compacted slots
    {
        uint32_t i = p_first;
        uint64_t prot_end = __ldg(a.starts + i + 1);
        uint64_t code = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) code = code * 40u + SIGK_SYM(j);
        int o = (int)(incl - mine);
        uint32_t run_i = 0, run_c = 0;              // occurrences of the protein I am inside
#pragma unroll
        for (int j = 0; j < ENC_PPT; ++j) {
            const uint64_t g = g_first + j;
            if (j > 0) code = (code - SIGK_SYM(j - 1) * SIGK_P7) * 40u + SIGK_SYM(j + 7);
            while (g >= prot_end && i + 1 < a.n_prot) { ++i; prot_end = __ldg(a.starts + i + 1); }
            if ((valid_mask >> j) & 1u) {
                const int slot = stage_slot(o++);
                sm.keys[warp][slot] = sigk_pack_key(code, (unsigned)(prot_end - g));
                sm.vals[warp][slot] = a.ordinal_base + i;
                if (i != run_i) {
                    if (run_c && a.prot_windows) atomicAdd(a.prot_windows + run_i, run_c);
                    run_i = i; run_c = 0;
                }
                ++run_c;
            }
        }
        if (run_c && a.prot_windows) atomicAdd(a.prot_windows + run_i, run_c);
    }
    __syncwarp();
    const uint64_t base = chained_scan_resolve_warp(scan_state, sub, total);
    if (last_sub && lane == 0) *n_out = base + total;
    for (uint32_t o = lane; o < total; o += 32) {
        const int slot = stage_slot((int)o);
        keys[base + o] = sm.keys[warp][slot];
        vals[base + o] = sm.vals[warp][slot];
    }
}

// ---- multi-GPU: encode and route in one pass ---------------------------------------------------
// The same window loop, but every record is written straight into the output region of the rank
// that owns its k-mer range (owner = number of splitter codes <= code), in canonical order inside
// each region: one chained scan per owner over the warp slices (a lane per owner walks back, as in
// the onesweep look-back).  This replaces encode + owner histogram + split pass.
// Per-owner 16-bit counters live four to a 64-bit word; the kernel is instantiated for 1, 2 or 4
// words (up to 4, 8, 16 ranks) so that the common small worlds keep a small register footprint.

struct EncSplitSmem {
    uint64_t keys[ENC_WARPS][ENC_STAGE];
    uint32_t vals[ENC_WARPS][ENC_STAGE];
    uint32_t obase[ENC_WARPS][16];          // first staging slot of each owner inside the warp
    uint64_t split[16];
    uint64_t *dkeys[16];
    uint32_t *dvals[16];
    uint32_t tile;
};

template <int WORDS>
SIGK_D uint32_t field16(const uint64_t (&w)[WORDS], uint32_t d) {
    uint64_t x = w[0];
#pragma unroll
    for (int k = 1; k < WORDS; ++k) x = (d >> 2) == (uint32_t)k ? w[k] : x;
    return (uint32_t)(x >> (16u * (d & 3u))) & 0xFFFFu;
}
template <int WORDS>
SIGK_D void bump16(uint64_t (&w)[WORDS], uint32_t d) {
    const uint64_t one = 1ull << (16u * (d & 3u));
    if (WORDS == 1) { w[0] += one; return; }
#pragma unroll
    for (int k = 0; k < WORDS; ++k) w[k] += (d >> 2) == (uint32_t)k ? one : 0ull;
}

template <int SPLIT_WORDS>
__global__ void __launch_bounds__(ENC_THREADS, 3)
encode_split_kernel(EncodeArgs a, const __grid_constant__ EncodeSplitArgs sp, uint32_t *__restrict__ ticket) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EncSplitSmem &sm = *reinterpret_cast<EncSplitSmem *>(smem_raw);

    __shared__ int8_t s_sym[256];
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t W = (uint32_t)sp.n_split + 1u;
    s_sym[tid] = (int8_t)sigk_symbol(tid);
    if (tid == 0) sm.tile = atomicAdd(ticket, 1u);
    if (tid < (unsigned)sp.n_split) sm.split[tid] = sp.split_codes[tid];
    if (tid < 16) { sm.dkeys[tid] = sp.dst_keys[tid]; sm.dvals[tid] = sp.dst_vals[tid]; }
    __syncthreads();
    const uint32_t sub = sm.tile * ENC_WARPS + warp;
    const uint64_t g0 = (uint64_t)sub * ENC_SUB;
    if (g0 >= a.total_res) return;
    const bool last_sub = g0 + ENC_SUB >= a.total_res;
    const uint32_t p_lo = __ldg(a.slice_prot + sub);
    const uint32_t p_hi = last_sub ? a.n_prot - 1 : __ldg(a.slice_prot + sub + 1);

    const uint64_t g_first = g0 + (uint64_t)lane * ENC_PPT;
    const uint4 w = ld_stream_u128(reinterpret_cast<const uint4 *>(a.res + g_first));
    uint32_t n0 = __shfl_down_sync(0xffffffffu, w.x, 1);
    uint32_t n1 = __shfl_down_sync(0xffffffffu, w.y, 1);
    if (lane == 31) {
        const uint2 nx = *reinterpret_cast<const uint2 *>(a.res + g_first + ENC_PPT);
        n0 = nx.x; n1 = nx.y;
    }
    const uint32_t words[6] = {w.x, w.y, w.z, w.w, n0, n1};
    // symbols stay packed four to a register (23 ints would not fit the register budget)
    uint32_t sw[6] = {0, 0, 0, 0, 0, 0};
    uint32_t bad = 0;
#pragma unroll
    for (int j = 0; j < ENC_PPT + 7; ++j) {
        const unsigned c = (words[j >> 2] >> (8 * (j & 3))) & 0xFFu;
        int sy = s_sym[c];                              // byte -> symbol 0..39, or -1 (table: 1/5 of the arithmetic form's instructions)
        if (sy < 0) { bad |= 1u << j; sy = 0; }
        sw[j >> 2] |= (uint32_t)sy << (8 * (j & 3));
    }
#define SIGK_SYM(j) ((uint64_t)((sw[(j) >> 2] >> (8 * ((j) & 3))) & 0xFFu))

    // pass A: valid windows, their owners (4 bits each), per-owner counts
    const uint32_t p_first = find_protein(a.starts, p_lo, p_hi, g_first < a.total_res ? g_first : a.total_res - 1);
    uint32_t valid_mask = 0;
    uint64_t owners = 0;
    uint64_t cnt[SPLIT_WORDS];
#pragma unroll
    for (int k = 0; k < SPLIT_WORDS; ++k) cnt[k] = 0;
    {
        uint32_t i = p_first;
        uint64_t prot_end = __ldg(a.starts + i + 1);
        uint64_t code = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) code = code * 40u + SIGK_SYM(j);
#pragma unroll
        for (int j = 0; j < ENC_PPT; ++j) {
            const uint64_t g = g_first + j;
            if (j > 0) code = (code - SIGK_SYM(j - 1) * SIGK_P7) * 40u + SIGK_SYM(j + 7);
            while (g >= prot_end && i + 1 < a.n_prot) { ++i; prot_end = __ldg(a.starts + i + 1); }
            if (((bad >> j) & 0xFFu) == 0 && g + SIGK_K_DEV <= prot_end) {
                valid_mask |= 1u << j;
                uint32_t d = 0;
                for (int k = 0; k < sp.n_split; ++k) d += code >= sm.split[k] ? 1u : 0u;
                owners |= (uint64_t)d << (4 * j);
                bump16(cnt, d);
            }
        }
    }
    // warp scan of the packed counters (a slice has at most 512 records: 16-bit fields never carry)
    uint64_t incl[SPLIT_WORDS];
#pragma unroll
    for (int k = 0; k < SPLIT_WORDS; ++k) {
        uint64_t v = cnt[k];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t y = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= (unsigned)o) v += y;
        }
        incl[k] = v;
    }
    uint64_t excl[SPLIT_WORDS], tot[SPLIT_WORDS];
#pragma unroll
    for (int k = 0; k < SPLIT_WORDS; ++k) { excl[k] = incl[k] - cnt[k]; tot[k] = __shfl_sync(0xffffffffu, incl[k], 31); }
    // lane d: this slice's record count for owner d; staging base of owner d; publish
    const uint32_t t_d = lane < W ? field16(tot, lane) : 0u;
    uint32_t ob = t_d;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, ob, o);
        if (lane >= (unsigned)o) ob += y;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, ob, 15);
    ob -= t_d;
    if (lane < 16) sm.obase[warp][lane] = ob;
    uint64_t *my_state = sp.owner_state + (size_t)sub * W + lane;
    if (lane < W) st_volatile_u64(my_state, (sub == 0 ? SIGK_CS_PRE : SIGK_CS_AGG) | (uint64_t)t_d);
    __syncwarp();

    // pass B: roll the code again, emit every valid window to its owner's part of the staging area
    // (the empty asm makes the symbols opaque so that pass A's partial products are not kept live)
#pragma unroll
    for (int k = 0; k < 6; ++k) asm volatile("" : "+r"(sw[k]));
    {
        uint32_t i = p_first;
        uint64_t prot_end = __ldg(a.starts + i + 1);
        uint64_t code = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) code = code * 40u + SIGK_SYM(j);
        uint64_t run[SPLIT_WORDS];
#pragma unroll
        for (int k = 0; k < SPLIT_WORDS; ++k) run[k] = excl[k];
        uint32_t run_i = 0, run_c = 0;
#pragma unroll
        for (int j = 0; j < ENC_PPT; ++j) {
            const uint64_t g = g_first + j;
            if (j > 0) code = (code - SIGK_SYM(j - 1) * SIGK_P7) * 40u + SIGK_SYM(j + 7);
            while (g >= prot_end && i + 1 < a.n_prot) { ++i; prot_end = __ldg(a.starts + i + 1); }
            if ((valid_mask >> j) & 1u) {
                const uint32_t d = (uint32_t)(owners >> (4 * j)) & 15u;
                const int slot = stage_slot((int)(sm.obase[warp][d] + field16(run, d)));
                bump16(run, d);
                sm.keys[warp][slot] = sigk_pack_key(code, (unsigned)(prot_end - g));
                sm.vals[warp][slot] = a.ordinal_base + i;
                if (i != run_i) {
                    if (run_c && a.prot_windows) atomicAdd(a.prot_windows + run_i, run_c);
                    run_i = i; run_c = 0;
                }
                ++run_c;
            }
        }
        if (run_c && a.prot_windows) atomicAdd(a.prot_windows + run_i, run_c);
    }

    // resolve: lane d walks back over the earlier slices of owner d
    uint64_t before = 0;
    if (lane < W) {
        if (sub > 0) {
            int64_t t = (int64_t)sub - 1;
            for (;;) {
                const uint64_t v = ld_volatile_u64(sp.owner_state + (size_t)t * W + lane);
                const uint64_t flag = v >> 62;
                if (flag == 0) continue;
                before += v & SIGK_CS_VAL;
                if (flag == 2) break;
                --t;
            }
            st_volatile_u64(my_state, SIGK_CS_PRE | (before + t_d));
        }
        if (before + t_d > sp.region_stride) atomicOr(sp.overflow, 1u);
        if (last_sub) sp.owner_totals[lane] = before + t_d;
    }
    __syncwarp();
    if (*reinterpret_cast<volatile uint32_t *>(sp.overflow)) return;      // regions too small: the caller falls back

    // One owner at a time: its records are a contiguous run of the staging area and go to a contiguous run
    // of the owner's region.  The body of every run leaves as 16-byte stores (two keys / four values per
    // lane) — what a peer GPU's region needs to be fed at NVLink rate; the unaligned ends go out singly.
    for (uint32_t d = 0; d < W; ++d) {
        const uint32_t len = __shfl_sync(0xffffffffu, t_d, d);
        if (len == 0) continue;
        const uint32_t s0 = __shfl_sync(0xffffffffu, ob, d);
        const uint64_t dst0 = __shfl_sync(0xffffffffu, before, d);
        uint64_t *dk = sm.dkeys[d] + dst0;
        uint32_t *dv = sm.dvals[d] + dst0;
        const uint32_t hk = (uint32_t)(dst0 & 1u);                        // keys before the first 16-byte boundary
        if (lane == 0 && hk) dk[0] = sm.keys[warp][stage_slot((int)s0)];
        const uint32_t pairs = (len - hk) >> 1;
        for (uint32_t q = lane; q < pairs; q += 32) {
            const uint32_t o = s0 + hk + 2 * q;
            ulonglong2 v;
            v.x = sm.keys[warp][stage_slot((int)o)];
            v.y = sm.keys[warp][stage_slot((int)o + 1)];
            *reinterpret_cast<ulonglong2 *>(dk + hk + 2 * q) = v;
        }
        if (lane == 0 && ((len - hk) & 1u)) dk[len - 1] = sm.keys[warp][stage_slot((int)(s0 + len - 1))];
        const uint32_t hv = min(len, (uint32_t)((4u - (uint32_t)(dst0 & 3u)) & 3u));
        if (lane < hv) dv[lane] = sm.vals[warp][stage_slot((int)(s0 + lane))];
        const uint32_t quads = (len - hv) >> 2;
        for (uint32_t q = lane; q < quads; q += 32) {
            const uint32_t o = s0 + hv + 4 * q;
            uint4 v;
            v.x = sm.vals[warp][stage_slot((int)o)];
            v.y = sm.vals[warp][stage_slot((int)o + 1)];
            v.z = sm.vals[warp][stage_slot((int)o + 2)];
            v.w = sm.vals[warp][stage_slot((int)o + 3)];
            *reinterpret_cast<uint4 *>(dv + hv + 4 * q) = v;
        }
        const uint32_t tv = (len - hv) & 3u;
        if (lane < tv) dv[len - tv + lane] = sm.vals[warp][stage_slot((int)(s0 + len - tv + lane))];
    }
}

// slice_prot[b] = protein holding position b * ENC_SUB: one binary search per slice, done
// once per upload instead of once per warp per build.
__global__ void slice_index_kernel(const uint64_t *__restrict__ starts, uint32_t n_prot, uint64_t total_res,
                                   uint32_t *__restrict__ slice_prot, uint64_t n_slices) {
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_slices) return;
    uint64_t g = b * ENC_SUB;
    if (g >= total_res) g = total_res - 1;
    slice_prot[b] = find_protein(starts, 0, n_prot - 1, g);
}

}  // namespace

cudaError_t launch_slice_index(const uint64_t *starts, uint32_t n_prot, uint64_t total_res, uint32_t *slice_prot,
                               cudaStream_t stream) {
    if (total_res == 0 || n_prot == 0) return cudaSuccess;
    const uint64_t n = encode_slices(total_res);
    slice_index_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(starts, n_prot, total_res, slice_prot, n);
    return cudaGetLastError();
}

cudaError_t launch_encode_split(const EncodeArgs &a, const EncodeSplitArgs &sp, uint32_t *ticket, cudaStream_t stream) {
    if (a.total_res == 0 || a.n_prot == 0) return cudaSuccess;
    if (sp.n_split < 1 || sp.n_split > 15) return cudaErrorInvalidValue;
    // 1, 2 or 4 counter words for up to 4, 8, 16 ranks.  SIGK_TEST_SPLIT_WORDS=2|4 forces a wider instantiation than the
    // rank count needs, so that a two-GPU box can exercise the kernels of the larger worlds (tests/multigpu_check.py).
    int words = sp.n_split < 4 ? 1 : sp.n_split < 8 ? 2 : 4;
    if (const char *force = std::getenv("SIGK_TEST_SPLIT_WORDS")) {
        const int v = std::atoi(force);
        if ((v == 2 || v == 4) && v > words) words = v;
    }
    auto kernel = words == 1 ? encode_split_kernel<1> : words == 2 ? encode_split_kernel<2> : encode_split_kernel<4>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EncSplitSmem));
    if (e != cudaSuccess) return e;
    kernel<<<(unsigned)encode_tiles(a.total_res), ENC_THREADS, sizeof(EncSplitSmem), stream>>>(a, sp, ticket);
    return cudaGetLastError();
}

cudaError_t launch_encode(const EncodeArgs &a, uint64_t *keys, uint32_t *vals, uint64_t *scan_state,
                          uint32_t *ticket, uint64_t *n_out, cudaStream_t stream) {
    if (a.total_res == 0 || a.n_prot == 0) return cudaSuccess;   // n_out stays 0
    // per device, so set on every launch (a process may drive several devices through several handles)
    cudaError_t attr = cudaFuncSetAttribute(encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EncSmem));
    if (attr != cudaSuccess) return attr;
    const uint64_t tiles = encode_tiles(a.total_res);
    encode_kernel<<<(unsigned)tiles, ENC_THREADS, sizeof(EncSmem), stream>>>(a, keys, vals, scan_state, ticket, n_out);
    return cudaGetLastError();
}

}  // namespace sigk
