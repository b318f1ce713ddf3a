// window_scan.cuh — the window loop of SignatureBuilder<K>::load_kmers_from_sequence
// (reference src/signature_build.tcc:162-180) as a per-lane device routine shared by the kernels
// that walk the residues: window_count_kernel, encode_split_kernel (encode.cu), encode_sort_kernel
// and encode_route_kernel (onesweep.cu).
//
//   - a window is valid iff all 8 bytes are in ok_prot_ (src/signature_build.h:102-103) and it lies
//     inside one protein (it < seq.end()-K+1);
//   - offset = (unsigned short)(len - p)   (:164);
//   - windows are visited in increasing position, proteins in input order = the multimap insertion
//     order of the reference's serial branch (:50-56).
//
// Layout: a warp owns one slice of WS_SUB = 512 consecutive residue positions, a lane 16 of them,
// read with one 128-bit load (512 contiguous bytes per warp); the 7 look-ahead residues come from
// the next lane by shuffle.  The arithmetic is 32-bit wherever the values allow it (the round-2
// profile of the first version: ~90 instructions per window, the 35-bit code recomputed from its
// eight symbols with 64-bit multiply-adds): the code of a window is two rolling base-20 codes of
// four residues joined by one wide multiply, positions are relative to the lane's first one, and
// the byte -> symbol table is read four residues to a register.
#pragma once

#include "kernels.h"
#include "sigk_common.cuh"

namespace sigk {

constexpr int WS_PPT = ENC_PPT;                 // window positions per lane
constexpr int WS_SUB = 32 * WS_PPT;             // positions per warp slice: 512
static_assert(WS_PPT == 16, "a lane reads its positions with one 128-bit load");

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
// the caller synchronises before the first ws_load.  Entry: rank 0..19 | lower case << 5, or 0x80
// for a byte outside ok_prot_.
SIGK_D void ws_fill_symbols(int8_t *s_sym) {
    if (threadIdx.x < 256) {
        const int sy = sigk_symbol(threadIdx.x);
        reinterpret_cast<uint8_t *>(s_sym)[threadIdx.x] = sy < 0 ? (uint8_t)0x80 : (uint8_t)sy;
    }
}

struct WindowLane {
    uint64_t g_first;       // first of the lane's 16 positions
    uint32_t rk[6];         // residue ranks 0..19, four to a word (16 own + 7 look-ahead; invalid bytes read 0)
    uint32_t bad;           // bit j: residue j is outside ok_prot_
    uint32_t low;           // bit j: residue j is lower case
    uint32_t p_first;       // protein holding g_first
};

// bits 0, 8, 16, 24 of x -> bits 0..3
SIGK_D uint32_t ws_gather4(uint32_t x) { return ((x & 0x01010101u) * 0x01020408u) >> 24; }

// Called by all 32 lanes of the warp that owns slice `sub` (sub * WS_SUB < total_res).
SIGK_D void ws_load(const EncodeArgs &a, const int8_t *s_sym_, uint32_t sub, WindowLane &w) {
    const uint8_t *s_sym = reinterpret_cast<const uint8_t *>(s_sym_);
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
    for (int k = 0; k < 6; ++k) {
        const uint32_t x = words[k];
        const uint32_t t = (uint32_t)s_sym[x & 0xFFu] | ((uint32_t)s_sym[(x >> 8) & 0xFFu] << 8) |
                           ((uint32_t)s_sym[(x >> 16) & 0xFFu] << 16) | ((uint32_t)s_sym[x >> 24] << 24);
        w.rk[k] = t & 0x1F1F1F1Fu;                       // (0x80 & 0x1F = 0: invalid residues read rank 0)
        w.low |= ws_gather4(t >> 5) << (4 * k);
        w.bad |= ws_gather4(t >> 7) << (4 * k);
    }
    w.p_first = find_protein(a.starts, p_lo, p_hi, w.g_first < a.total_res ? w.g_first : a.total_res - 1);
}

#define SIGK_WS_RANK(w, j) (((w).rk[(j) >> 2] >> (8 * ((j) & 3))) & 0xFFu)

// positions of a protein's end relative to the lane's first position, clamped so that they fit 32 bits with room to spare
SIGK_D uint32_t ws_rel(uint64_t prot_end, uint64_t g_first) {
    if (prot_end <= g_first) return 0u;                  // (positions past the last residue: the walk below moves on, nothing is valid)
    const uint64_t d = prot_end - g_first;
    return d > 0x40000000ull ? 0x40000000u : (uint32_t)d;
}

// bit j: the lane's window j is valid
SIGK_D uint32_t ws_valid_mask(const EncodeArgs &a, const WindowLane &w) {
    uint32_t valid = 0;
    uint32_t i = w.p_first;
    uint32_t rel_end = ws_rel(__ldg(a.starts + i + 1), w.g_first);
#pragma unroll
    for (int j = 0; j < WS_PPT; ++j) {
        while ((uint32_t)j >= rel_end && i + 1 < a.n_prot) { ++i; rel_end = ws_rel(__ldg(a.starts + i + 1), w.g_first); }
        if (((w.bad >> j) & 0xFFu) == 0 && (uint32_t)j + SIGK_K_DEV <= rel_end) valid |= 1u << j;
    }
    return valid;
}

// f(j, code35, mask8, offset16, protein) for every window j set in `valid`, in position order (protein = local index of
// the window's protein); per_protein(protein, n) once for every protein that has n > 0 of those windows.
// The lane's 16 positions are cut at the protein boundaries inside them (usually none): one trip of the outer loop
// per protein, the unrolled window loop predicated on the protein's share — no boundary test per window.
template <class F, class P>
SIGK_D void ws_for_each_code(const EncodeArgs &a, const WindowLane &w, uint32_t valid, F &&f, P &&per_protein) {
    if (valid == 0) return;
    uint32_t i = w.p_first, seg_lo = 0;
    const uint32_t g_lo = (uint32_t)w.g_first;
    for (;;) {
        const uint64_t prot_end = __ldg(a.starts + i + 1);
        const uint32_t rel_end = ws_rel(prot_end, w.g_first);
        const uint32_t seg_hi = rel_end < (uint32_t)WS_PPT ? rel_end : (uint32_t)WS_PPT;
        // windows that start in [seg_lo, seg_hi) start inside protein i (a valid one also ends inside it)
        const uint32_t m = seg_hi > seg_lo ? valid & ((1u << seg_hi) - 1u) & ~((1u << seg_lo) - 1u) : 0u;
        if (m) {
            per_protein(i, (uint32_t)__popc(m));
            const uint32_t off_base = (uint32_t)prot_end - g_lo;     // offsets need the low 16 bits only
            // h(t) = base-20 value of residues t..t+3 (rolling, < 160000); the window at j has code h(j) * 20^4 + h(j+4)
            uint32_t h = ((SIGK_WS_RANK(w, 0) * 20u + SIGK_WS_RANK(w, 1)) * 20u + SIGK_WS_RANK(w, 2)) * 20u + SIGK_WS_RANK(w, 3);
            uint32_t hq[4];                                          // h(t-4) .. h(t-1)
#pragma unroll
            for (int t = 0; t < WS_PPT + 4; ++t) {
                if (t > 0) h = (h - SIGK_WS_RANK(w, t - 1) * 8000u) * 20u + SIGK_WS_RANK(w, t + 3);
                if (t >= 4) {
                    const int j = t - 4;
                    if ((m >> j) & 1u) f(j, (uint64_t)hq[j & 3] * 160000u + h, (w.low >> j) & 0xFFu, (off_base - (uint32_t)j) & 0xFFFFu, i);
                }
                hq[t & 3] = h;
            }
        }
        if (rel_end >= (uint32_t)WS_PPT || i + 1 >= a.n_prot) break;
        seg_lo = seg_hi;
        ++i;
    }
}
template <class F>
SIGK_D void ws_for_each_code(const EncodeArgs &a, const WindowLane &w, uint32_t valid, F &&f) {
    ws_for_each_code(a, w, valid, f, [](uint32_t, uint32_t) {});
}

// f(j, key, protein): the same with the packed 12-byte record's key = code35 << 29 | mask8 << 21 | offset16 << 5
template <class F, class P>
SIGK_D void ws_for_each(const EncodeArgs &a, const WindowLane &w, uint32_t valid, F &&f, P &&per_protein) {
    ws_for_each_code(a, w, valid, [&](int j, uint64_t code, uint32_t mask, uint32_t off, uint32_t i) { f(j, sigk_pack_key(code, mask, off), i); },
                     per_protein);
}
template <class F>
SIGK_D void ws_for_each(const EncodeArgs &a, const WindowLane &w, uint32_t valid, F &&f) {
    ws_for_each(a, w, valid, f, [](uint32_t, uint32_t) {});
}

}  // namespace sigk
