// window_scan.cuh — the window loop of SignatureBuilder<K>::load_kmers_from_sequence
// (reference src/signature_build.tcc:162-180) as a per-lane device routine shared by the kernels
// that walk the residues: window_count_kernel, encode_split_kernel (encode.cu) and
// encode_sort_kernel (onesweep.cu).
//
//   - a window is valid iff all 8 bytes are in ok_prot_ (src/signature_build.h:102-103) and it lies
//     inside one protein (it < seq.end()-K+1);
//   - offset = (unsigned short)(len - p)   (:164);
//   - windows are visited in increasing position, proteins in input order = the multimap insertion
//     order of the reference's serial branch (:50-56).
//
// Layout: a warp owns one slice of WS_SUB = 512 consecutive residue positions, a lane 16 of them,
// read with one 128-bit load (512 contiguous bytes per warp); the 7 look-ahead residues come from
// the next lane by shuffle.  The case-folded base-20 code is rolled (one multiply-add per window),
// the case mask of a window is a bit field of the lane's lower-case flags.
#pragma once

#include "kernels.h"
#include "sigk_common.cuh"

namespace sigk {

constexpr int WS_PPT = ENC_PPT;                 // window positions per lane
constexpr int WS_SUB = 32 * WS_PPT;             // positions per warp slice: 512

// largest i in [lo, hi] with starts[i] <= g   (starts[lo] <= g is guaranteed)
SIGK_D uint32_t find_protein(const uint64_t *__restrict__ starts, uint32_t lo, uint32_t hi, uint64_t g) {
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo + 1) >> 1);
        if (__ldg(starts + mid) <= g) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

// byte -> symbol table in shared memory (256 entries), filled by the first 256 threads of the CTA;
// the caller synchronises before the first ws_load
SIGK_D void ws_fill_symbols(int8_t *s_sym) {
    if (threadIdx.x < 256) s_sym[threadIdx.x] = (int8_t)sigk_symbol(threadIdx.x);
}

struct WindowLane {
    uint64_t g_first;       // first of the lane's 16 positions
    uint32_t rk[6];         // residue ranks 0..19, four to a word (16 own + 7 look-ahead; invalid bytes read 0)
    uint32_t bad;           // bit j: residue j is outside ok_prot_
    uint32_t low;           // bit j: residue j is lower case
    uint32_t p_first;       // protein holding g_first
};

// Called by all 32 lanes of the warp that owns slice `sub` (sub * WS_SUB < total_res).
SIGK_D void ws_load(const EncodeArgs &a, const int8_t *s_sym, uint32_t sub, WindowLane &w) {
    const unsigned lane = threadIdx.x & 31u;
    const uint64_t g0 = (uint64_t)sub * WS_SUB;
    const bool last_sub = g0 + WS_SUB >= a.total_res;
    // protein range of the warp's positions, from the per-slice index built by slice_index_kernel
    const uint32_t p_lo = __ldg(a.slice_prot + sub);
    const uint32_t p_hi = last_sub ? a.n_prot - 1 : __ldg(a.slice_prot + sub + 1);
    w.g_first = g0 + (uint64_t)lane * WS_PPT;
    const uint4 q = ld_stream_u128(reinterpret_cast<const uint4 *>(a.res + w.g_first));
    uint32_t n0 = __shfl_down_sync(0xffffffffu, q.x, 1);
    uint32_t n1 = __shfl_down_sync(0xffffffffu, q.y, 1);
    if (lane == 31) {
        const uint2 nx = *reinterpret_cast<const uint2 *>(a.res + w.g_first + WS_PPT);   // the buffer is padded (ENC_PAD)
        n0 = nx.x; n1 = nx.y;
    }
    const uint32_t words[6] = {q.x, q.y, q.z, q.w, n0, n1};
    w.bad = 0; w.low = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) w.rk[k] = 0;
#pragma unroll
    for (int j = 0; j < WS_PPT + 7; ++j) {
        const unsigned c = (words[j >> 2] >> (8 * (j & 3))) & 0xFFu;
        int sy = s_sym[c];                               // rank | lower << 5, or -1
        if (sy < 0) { w.bad |= 1u << j; sy = 0; }
        w.rk[j >> 2] |= (uint32_t)(sy & 31) << (8 * (j & 3));
        w.low |= (uint32_t)(sy >> 5) << j;
    }
    w.p_first = find_protein(a.starts, p_lo, p_hi, w.g_first < a.total_res ? w.g_first : a.total_res - 1);
}

#define SIGK_WS_RANK(w, j) ((uint64_t)(((w).rk[(j) >> 2] >> (8 * ((j) & 3))) & 0xFFu))

// bit j: the lane's window j is valid
SIGK_D uint32_t ws_valid_mask(const EncodeArgs &a, const WindowLane &w) {
    uint32_t valid = 0;
    uint32_t i = w.p_first;
    uint64_t prot_end = __ldg(a.starts + i + 1);
#pragma unroll
    for (int j = 0; j < WS_PPT; ++j) {
        const uint64_t g = w.g_first + j;
        while (g >= prot_end && i + 1 < a.n_prot) { ++i; prot_end = __ldg(a.starts + i + 1); }
        if (((w.bad >> j) & 0xFFu) == 0 && g + SIGK_K_DEV <= prot_end) valid |= 1u << j;
    }
    return valid;
}

// f(j, key, protein) for every window j set in `valid`, in position order.
// key = code35 << 29 | mask8 << 21 | offset16 << 5; protein = local index of the window's protein.
template <class F>
SIGK_D void ws_for_each(const EncodeArgs &a, const WindowLane &w, uint32_t valid, F &&f) {
    uint32_t i = w.p_first;
    uint64_t prot_end = __ldg(a.starts + i + 1);
    uint64_t code = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) code = code * 20u + SIGK_WS_RANK(w, j);
#pragma unroll
    for (int j = 0; j < WS_PPT; ++j) {
        const uint64_t g = w.g_first + j;
        if (j > 0) code = (code - SIGK_WS_RANK(w, j - 1) * SIGK_P7) * 20u + SIGK_WS_RANK(w, j + 7);
        while (g >= prot_end && i + 1 < a.n_prot) { ++i; prot_end = __ldg(a.starts + i + 1); }
        if ((valid >> j) & 1u) f(j, sigk_pack_key(code, (w.low >> j) & 0xFFu, (unsigned)(prot_end - g)), i);
    }
}

}  // namespace sigk
