// encode.cu — stage 1: slide the 8-residue window over the packed protein
// buffer, drop windows the reference drops, and either count (window_count_kernel)
// or emit (encode_split_kernel) one 12-byte record per valid window in canonical
// (insertion) order.  The window loop itself is window_scan.cuh.
//
// Replaces the window loop of SignatureBuilder<K>::load_kmers_from_sequence
// (reference src/signature_build.tcc:162-180).
//
//   window_count_kernel   the digit histograms of every sort pass, taken from the residues before a
//                         single record exists (0.6 GB read instead of the 4.7 GB key re-read a
//                         histogram of encoded records costs), the number of records with a
//                         lower-case residue (the side run), and every protein's window count.
//                         It is what lets the first radix pass be fused with the encode
//                         (encode_sort_kernel, onesweep.cu).
//   encode_split_kernel   encode + route: every record goes straight to the output region of the rank
//                         that owns its k-mer range (multi-GPU; with one owner it is the plain encode).
#include "kernels.h"
#include "sigk_common.cuh"
#include "window_scan.cuh"

#include <cstdlib>

namespace sigk {

namespace {

constexpr int ENC_WARPS = ENC_THREADS / 32;
constexpr int ENC_SUB = WS_SUB;                       // window positions one warp owns: 512
// compacted records are staged with one pad slot per 16 so that the
// thread-contiguous writes (stride ~16 records between lanes) spread over banks
constexpr int ENC_STAGE = ENC_SUB + ENC_SUB / 16;
SIGK_D int stage_slot(int o) { return o + (o >> 4); }

// ---- count pass -----------------------------------------------------------------------------
// A persistent grid walks the slices; digit counts go to shared-memory counters (one flush per
// CTA).  Records whose window has a lower-case residue (mask8 != 0) are counted into the side bin
// of pass 0 only: the first sort pass diverts them into the side run, which gets its own small
// histogram later (launch_histogram over the side run).
constexpr int WC_THREADS = 256;
constexpr int WC_WARPS = WC_THREADS / 32;

template <int NPASS, bool STD>
__global__ void __launch_bounds__(WC_THREADS, 3)
window_count_kernel(EncodeArgs a, const __grid_constant__ PassPlan plan, uint64_t *__restrict__ hist, uint32_t n_slices) {
    __shared__ uint32_t sh[NPASS * SIGK_RADIX];
    __shared__ int8_t s_sym[256];
    __shared__ uint32_t s_side;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    ws_fill_symbols(s_sym);
    for (int j = tid; j < NPASS * SIGK_RADIX; j += WC_THREADS) sh[j] = 0;
    if (tid == 0) s_side = 0;
    __syncthreads();
    uint32_t side = 0;
    for (uint64_t sub = (uint64_t)blockIdx.x * WC_WARPS + warp; sub < n_slices; sub += (uint64_t)gridDim.x * WC_WARPS) {
        if (sub * WS_SUB >= a.total_res) break;
        WindowLane w;
        ws_load(a, s_sym, (uint32_t)sub, w);
        const uint32_t valid = ws_valid_mask(a, w);
        ws_for_each_code(a, w, valid, [&](int, uint64_t code, uint32_t mask, uint32_t, uint32_t) {
            if (mask == 0) {
                if (STD) {
                    // the standard plan (9/9/9/8 bits from key bit 29 = code bit 0): the digits with 32-bit operations
                    const uint32_t lo = (uint32_t)code, hi = (uint32_t)(code >> 32);
                    atomicAdd(&sh[0 * SIGK_RADIX + (lo & 511u)], 1u);
                    atomicAdd(&sh[1 * SIGK_RADIX + ((lo >> 9) & 511u)], 1u);
                    atomicAdd(&sh[2 * SIGK_RADIX + ((lo >> 18) & 511u)], 1u);
                    atomicAdd(&sh[3 * SIGK_RADIX + ((lo >> 27) | (hi << 5))], 1u);
                } else {
                    // (the digits of the key's bit fields, taken from the code: key = code << 29 | ...)
#pragma unroll
                    for (int p = 0; p < NPASS; ++p)
                        atomicAdd(&sh[p * SIGK_RADIX + ((uint32_t)(code >> (plan.lo[p] - SIGK_KEY_CODE35_SHIFT)) & ((1u << plan.bits[p]) - 1u))], 1u);
                }
            } else {
                ++side;
            }
        }, [&](uint32_t i, uint32_t n_windows) { if (a.prot_windows) atomicAdd(a.prot_windows + i, n_windows); });
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) side += __shfl_xor_sync(0xffffffffu, side, o);
    if (lane == 0 && side) atomicAdd(&s_side, side);
    __syncthreads();
    for (int j = tid; j < NPASS * SIGK_RADIX; j += WC_THREADS)
        if (sh[j]) atomicAdd(reinterpret_cast<unsigned long long *>(hist + (size_t)(j / SIGK_RADIX) * SIGK_BINS + (j % SIGK_RADIX)),
                             (unsigned long long)sh[j]);
    if (tid == 0 && s_side) atomicAdd(reinterpret_cast<unsigned long long *>(hist + SIGK_SIDE_BIN), (unsigned long long)s_side);
}

// ---- encode and route in one pass -----------------------------------------------------------
// Every record is written straight into the output region of the rank that owns its k-mer range
// (owner = number of splitter codes <= the record's case-folded code35), in canonical order inside
// each region: one chained scan per owner over the warp slices (a lane per owner walks back, as in
// the onesweep look-back).  With no splitters there is one owner and the kernel is the plain encode.
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
    static_assert(ENC_THREADS == 256, "one symbol table entry per thread");
    ws_fill_symbols(s_sym);
    if (tid == 0) sm.tile = atomicAdd(ticket, 1u);
    if (tid < (unsigned)sp.n_split) sm.split[tid] = sp.split_codes[tid];
    if (tid < 16) { sm.dkeys[tid] = sp.dst_keys[tid]; sm.dvals[tid] = sp.dst_vals[tid]; }
    __syncthreads();
    const uint32_t sub = sm.tile * ENC_WARPS + warp;
    const uint64_t g0 = (uint64_t)sub * ENC_SUB;
    if (g0 >= a.total_res) return;
    const bool last_sub = g0 + ENC_SUB >= a.total_res;

    WindowLane w;
    ws_load(a, s_sym, sub, w);
    const uint32_t valid_mask = ws_valid_mask(a, w);

    // pass A: the owners of the valid windows (4 bits each), per-owner counts
    uint64_t owners = 0;
    uint64_t cnt[SPLIT_WORDS];
#pragma unroll
    for (int k = 0; k < SPLIT_WORDS; ++k) cnt[k] = 0;
    if (sp.n_split > 0) {
        ws_for_each(a, w, valid_mask, [&](int j, uint64_t key, uint32_t) {
            const uint64_t code = sigk_key_code35(key);
            uint32_t d = 0;
            for (int k = 0; k < sp.n_split; ++k) d += code >= sm.split[k] ? 1u : 0u;
            owners |= (uint64_t)d << (4 * j);
            bump16(cnt, d);
        });
    } else {
        cnt[0] = (uint64_t)__popc(valid_mask);
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
    ob -= t_d;
    if (lane < 16) sm.obase[warp][lane] = ob;
    uint64_t *my_state = sp.owner_state + (size_t)sub * W + lane;
    if (lane < W) st_volatile_u64(my_state, (sub == 0 ? SIGK_CS_PRE : SIGK_CS_AGG) | (uint64_t)t_d);
    __syncwarp();

    // pass B: roll the code again, emit every valid window to its owner's part of the staging area
    // (the empty asm makes the symbols opaque so that pass A's partial products are not kept live)
#pragma unroll
    for (int k = 0; k < 6; ++k) asm volatile("" : "+r"(w.rk[k]));
    {
        uint64_t run[SPLIT_WORDS];
#pragma unroll
        for (int k = 0; k < SPLIT_WORDS; ++k) run[k] = excl[k];
        ws_for_each(a, w, valid_mask, [&](int j, uint64_t key, uint32_t i) {
            const uint32_t d = (uint32_t)(owners >> (4 * j)) & 15u;
            const int slot = stage_slot((int)(sm.obase[warp][d] + field16(run, d)));
            bump16(run, d);
            sm.keys[warp][slot] = key;
            sm.vals[warp][slot] = a.ordinal_base + i;
        }, [&](uint32_t i, uint32_t n_windows) { if (a.prot_windows) atomicAdd(a.prot_windows + i, n_windows); });
    }

    // resolve: lane d walks back over the earlier slices of owner d
    uint64_t before = 0;
    if (lane < W) {
        if (sub > 0) {
            int64_t t = (int64_t)sub - 1;
            for (;;) {
                const uint64_t v = ld_volatile_u64(sp.owner_state + (size_t)t * W + lane);
                const uint64_t flag = v >> 62;
                if (flag == 0) { __nanosleep(20); continue; }
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
    if (*reinterpret_cast<volatile uint32_t *>(sp.overflow)) return;      // regions too small: the caller grows them and retries

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

cudaError_t launch_window_count(const EncodeArgs &a, const PassPlan &plan, uint64_t *hist, int sm_count, cudaStream_t stream) {
    if (a.total_res == 0 || a.n_prot == 0) return cudaSuccess;
    const uint64_t n_slices = (a.total_res + WS_SUB - 1) / WS_SUB;
    const uint64_t want = (n_slices + WC_WARPS - 1) / WC_WARPS;
    const unsigned grid = (unsigned)std::min<uint64_t>(want, (uint64_t)sm_count * 8);
    for (int p = 0; p < plan.npass; ++p) if (plan.lo[p] < SIGK_KEY_CODE35_SHIFT) return cudaErrorInvalidValue;      // the passes of the main run sort code bits only
    const bool std_plan = plan.npass == 4 && plan.lo[0] == 29 && plan.bits[0] == 9 && plan.bits[1] == 9 && plan.bits[2] == 9 && plan.bits[3] == 8;
    if (std_plan) { window_count_kernel<4, true><<<grid, WC_THREADS, 0, stream>>>(a, plan, hist, (uint32_t)n_slices); return cudaGetLastError(); }
    switch (plan.npass) {
        case 3: window_count_kernel<3, false><<<grid, WC_THREADS, 0, stream>>>(a, plan, hist, (uint32_t)n_slices); break;
        case 4: window_count_kernel<4, false><<<grid, WC_THREADS, 0, stream>>>(a, plan, hist, (uint32_t)n_slices); break;
        case 5: window_count_kernel<5, false><<<grid, WC_THREADS, 0, stream>>>(a, plan, hist, (uint32_t)n_slices); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_encode_split(const EncodeArgs &a, const EncodeSplitArgs &sp, uint32_t *ticket, cudaStream_t stream) {
    if (a.total_res == 0 || a.n_prot == 0) return cudaSuccess;
    if (sp.n_split < 0 || sp.n_split > 15) return cudaErrorInvalidValue;
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

}  // namespace sigk
