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

    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) sm.tile = atomicAdd(ticket, 1u);
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

    int s[ENC_PPT + 7];
    uint32_t bad = 0;
#pragma unroll
    for (int j = 0; j < ENC_PPT + 7; ++j) {
        const unsigned c = (words[j >> 2] >> (8 * (j & 3))) & 0xFFu;
        int sy = sigk_symbol(c);
        if (sy < 0) { bad |= 1u << j; sy = 0; }
        s[j] = sy;
    }

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
        for (int j = 0; j < 8; ++j) code = code * 40u + (uint64_t)s[j];
        int o = (int)(incl - mine);
        uint32_t run_i = 0, run_c = 0;              // occurrences of the protein I am inside
#pragma unroll
        for (int j = 0; j < ENC_PPT; ++j) {
            const uint64_t g = g_first + j;
            if (j > 0) code = (code - (uint64_t)s[j - 1] * SIGK_P7) * 40u + (uint64_t)s[j + 7];
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

cudaError_t launch_encode(const EncodeArgs &a, uint64_t *keys, uint32_t *vals, uint64_t *scan_state,
                          uint32_t *ticket, uint64_t *n_out, cudaStream_t stream) {
    if (a.total_res == 0 || a.n_prot == 0) return cudaSuccess;   // n_out stays 0
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EncSmem));
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const uint64_t tiles = encode_tiles(a.total_res);
    encode_kernel<<<(unsigned)tiles, ENC_THREADS, sizeof(EncSmem), stream>>>(a, keys, vals, scan_state, ticket, n_out);
    return cudaGetLastError();
}

}  // namespace sigk
