// sigk_common.cuh — shared definitions for the libsigk kernels (sm_100a).
//
// Record layout in HBM (struct-of-arrays, 12 bytes per k-mer occurrence):
//   key  u64 = code35 << 29 | mask8 << 21 | offset16 << 5      (low 5 bits zero)
//   val  u32 = protein ordinal (index into the packed input, canonical order)
// code35 is the base-20 value of the 8 residues with their case folded away (the 20 amino-acid
// letters of src/signature_build.h:102-103 ranked in ASCII order; 20^8 < 2^35), mask8 says which of
// the 8 residues are lower case (first residue = bit 0).  The reference keeps case ('a' != 'A',
// src/signature_build.tcc:167), so a k-mer is the pair (code35, mask8) = the 43-bit "group code"
// key >> 21.  Nearly every window of real proteins is all upper case (mask8 == 0): those records
// are sorted on the 35 code bits alone (four 9/9/9/8-bit passes instead of the five a 43-bit
// byte-order code needs); the rare records with a lower-case residue are diverted by the first
// pass into a side run that is sorted on all 43 bits.  Table order is therefore: k-mers without a
// lower-case residue first, in byte order; then the others by (case-folded bytes, mask8).
// offset16 is the reference's `unsigned short n = distance(it, seq.end())`
// (src/signature_build.tcc:164).  Everything else KmerAttributes carries
// (func_index, seq_id, protein_length; src/kmer_data.h:105-112) is looked up by
// ordinal in L2-resident per-protein arrays.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#define SIGK_CODE35_BITS 35
#define SIGK_MASK_BITS 8
#define SIGK_CODE_BITS 43                  // group code = code35 << 8 | mask8
#define SIGK_KEY_CODE_SHIFT 21             // group code = key >> 21
#define SIGK_KEY_CODE35_SHIFT 29           // code35 = key >> 29
#define SIGK_KEY_OFFSET_SHIFT 5
#define SIGK_K_DEV 8
#ifndef SIGK_RADIX_BITS
#define SIGK_RADIX_BITS 9         // 35 code bits in 4 passes (9,9,9,8)
#endif
#define SIGK_RADIX (1 << SIGK_RADIX_BITS)
// digit counters / bin bases / look-back rows are padded to this many entries: 512 digits, the
// side bin of the first pass (SIGK_SIDE_BIN: records with a lower-case residue), a total
#define SIGK_BINS (SIGK_RADIX + 8)
#define SIGK_SIDE_BIN SIGK_RADIX

#define SIGK_HD __host__ __device__ __forceinline__
#define SIGK_D __device__ __forceinline__

// Letters A C D E F G H I K L M N P Q R S T V W Y as a bit set over 'A'+bit.
#define SIGK_AA_MASK 0x016FBDFDu

// Symbol of a residue byte: rank 0..19 of the letter among the 20 amino acids (ASCII order), plus 32
// when it is lower case; -1 for anything outside ok_prot_ (B J O U X Z * digits ...).
SIGK_HD int sigk_symbol(unsigned c) {
    const unsigned idx = (c & 0xDFu) - 0x41u;
    const bool letter = ((c & 0xC0u) == 0x40u) && idx < 26u && ((SIGK_AA_MASK >> idx) & 1u);
    if (!letter) return -1;
#ifdef __CUDA_ARCH__
    const int rank = __popc(SIGK_AA_MASK & ((1u << idx) - 1u));
#else
    const int rank = __builtin_popcount(SIGK_AA_MASK & ((1u << idx) - 1u));
#endif
    return rank + ((c & 0x20u) ? 32 : 0);
}

// 20^7
#define SIGK_P7 1280000000ULL

SIGK_HD uint64_t sigk_pack_key(uint64_t code35, unsigned mask8, unsigned offset16) {
    return (code35 << SIGK_KEY_CODE35_SHIFT) | ((uint64_t)(mask8 & 0xFFu) << SIGK_KEY_CODE_SHIFT) |
           ((uint64_t)(offset16 & 0xFFFFu) << SIGK_KEY_OFFSET_SHIFT);
}
SIGK_HD uint64_t sigk_key_code(uint64_t key) { return key >> SIGK_KEY_CODE_SHIFT; }            // 43-bit group code
SIGK_HD uint64_t sigk_key_code35(uint64_t key) { return key >> SIGK_KEY_CODE35_SHIFT; }
SIGK_HD unsigned sigk_key_mask(uint64_t key) { return (unsigned)(key >> SIGK_KEY_CODE_SHIFT) & 0xFFu; }
SIGK_HD unsigned sigk_key_offset(uint64_t key) { return (unsigned)(key >> SIGK_KEY_OFFSET_SHIFT) & 0xFFFFu; }

// upper-case residue letter of rank 0..19
SIGK_HD unsigned sigk_rank_ascii(unsigned u) {
#ifdef __CUDA_ARCH__
    // "ACDEFGHIKLMNPQRSTVWY" as little-endian words; byte select by __byte_perm
    unsigned c;
    if (u < 8u) c = __byte_perm(0x45444341u, 0x49484746u, u);             // A C D E | F G H I
    else if (u < 16u) c = __byte_perm(0x4E4D4C4Bu, 0x53525150u, u - 8u);   // K L M N | P Q R S
    else c = __byte_perm(0x59575654u, 0u, u - 16u);                        // T V W Y
    return c & 0xFFu;
#else
    return (unsigned char)"ACDEFGHIKLMNPQRSTVWY"[u];
#endif
}

// group code -> 8 ASCII bytes packed little-endian (first residue in the low byte), ready for one
// 8-byte store into the kmer column.  One 64-bit division splits code35 into two base-20 halves;
// the rest is 32-bit arithmetic; the case mask ORs 0x20 into the lower-case positions.
SIGK_HD uint64_t sigk_spread_mask(unsigned mask8) {
    // bit j of mask8 -> 0x20 in byte j
    uint64_t m = 0;
    for (int j = 0; j < 8; ++j) m |= (uint64_t)((mask8 >> j) & 1u) << (8 * j + 5);
    return m;
}
SIGK_HD uint64_t sigk_code_to_ascii(uint64_t gcode) {
    const uint64_t code35 = gcode >> SIGK_MASK_BITS;
    const uint32_t hi = (uint32_t)(code35 / 160000ull);                     // 20^4
    const uint32_t lo = (uint32_t)(code35 - (uint64_t)hi * 160000ull);
    uint32_t w[2];
    uint32_t v[2] = {hi, lo};
    for (int h = 0; h < 2; ++h) {
        uint32_t x = v[h];
        const uint32_t s3 = x % 20u; x /= 20u;
        const uint32_t s2 = x % 20u; x /= 20u;
        const uint32_t s1 = x % 20u; x /= 20u;
        w[h] = sigk_rank_ascii(x) | (sigk_rank_ascii(s1) << 8) | (sigk_rank_ascii(s2) << 16) | (sigk_rank_ascii(s3) << 24);
    }
    return ((uint64_t)w[0] | ((uint64_t)w[1] << 32)) | sigk_spread_mask((unsigned)gcode & 0xFFu);
}

#ifdef __CUDACC__

SIGK_D unsigned lane_id() { return threadIdx.x & 31u; }

SIGK_D uint64_t ld_volatile_u64(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
SIGK_D void st_volatile_u64(uint64_t *p, uint64_t v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
SIGK_D uint32_t ld_volatile_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
SIGK_D void st_volatile_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Streaming (read-once) loads: keep them out of L1.
SIGK_D uint64_t ld_stream_u64(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
SIGK_D uint32_t ld_stream_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
SIGK_D uint4 ld_stream_u128(const uint4 *p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}

// Gathers from the per-protein table: marked evict-last in L2 so that the records streaming past (each read
// once) do not push the table out — with several ranks' proteins in it the table is tens of megabytes and
// every record of the reduce stage looks one entry up.
SIGK_D uint64_t l2_evict_last_policy() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
SIGK_D uint32_t ld_keep_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(l2_evict_last_policy()));
    return v;
}
SIGK_D uint2 ld_keep_u32x2(const uint2 *p) {
    uint2 v;
    asm volatile("ld.global.nc.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(l2_evict_last_policy()));
    return v;
}

// Loads of data that is read once (the sorted records in the reduce stage): L1 no-allocate and evict-first in L2,
// the counterpart of ld_keep_* — the stream then recycles its own L2 lines instead of the table's.
SIGK_D uint64_t l2_evict_first_policy() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
SIGK_D uint64_t ld_once_u64(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(l2_evict_first_policy()));
    return v;
}
SIGK_D uint32_t ld_once_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(l2_evict_first_policy()));
    return v;
}
SIGK_D ulonglong2 ld_once_u64x2(const ulonglong2 *p) {
    ulonglong2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;" : "=l"(v.x), "=l"(v.y) : "l"(p), "l"(l2_evict_first_policy()));
    return v;
}
SIGK_D uint4 ld_once_u128(const uint4 *p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p), "l"(l2_evict_first_policy()));
    return v;
}

// ---- single-value chained scan (decoupled look-back) ------------------------
// state[t] = flag << 62 | value.  flag 0 = not ready, 1 = tile aggregate,
// 2 = inclusive prefix.  Tiles are tickets taken in launch order, so every
// predecessor of a running tile is running or done and publishes its aggregate
// before it waits on anything: the spin below always terminates.
#define SIGK_CS_AGG (1ULL << 62)
#define SIGK_CS_PRE (2ULL << 62)
#define SIGK_CS_VAL ((1ULL << 62) - 1)

// Called by ONE thread of the tile.  publish, then (after any independent work) resolve.
SIGK_D void chained_scan_publish(uint64_t *state, uint32_t tile, uint64_t aggregate) {
    st_volatile_u64(state + tile, (tile == 0 ? SIGK_CS_PRE : SIGK_CS_AGG) | aggregate);
}
SIGK_D uint64_t chained_scan_resolve(uint64_t *state, uint32_t tile, uint64_t aggregate) {
    if (tile == 0) return 0;
    uint64_t excl = 0;
    int64_t t = (int64_t)tile - 1;
    for (;;) {
        const uint64_t v = ld_volatile_u64(state + t);
        const uint64_t flag = v >> 62;
        if (flag == 0) continue;
        excl += v & SIGK_CS_VAL;
        if (flag == 2) break;
        --t;
    }
    st_volatile_u64(state + tile, SIGK_CS_PRE | (excl + aggregate));
    return excl;
}
// Returns the exclusive prefix of `aggregate`.
SIGK_D uint64_t chained_scan_exclusive(uint64_t *state, uint32_t tile, uint64_t aggregate) {
    chained_scan_publish(state, tile, aggregate);
    return chained_scan_resolve(state, tile, aggregate);
}

// The same scan, called by a FULL WARP for its own tile: 32 predecessors are
// inspected per round trip instead of one.  Split in two so that the caller can
// put independent work between publishing its aggregate and needing the prefix
// (by then the predecessors have usually published too).
SIGK_D void chained_scan_publish_warp(uint64_t *state, uint32_t tile, uint64_t aggregate) {
    if ((threadIdx.x & 31u) == 0) st_volatile_u64(state + tile, (tile == 0 ? SIGK_CS_PRE : SIGK_CS_AGG) | aggregate);
}
// All lanes return the exclusive prefix.
SIGK_D uint64_t chained_scan_resolve_warp(uint64_t *state, uint32_t tile, uint64_t aggregate) {
    const unsigned lane = threadIdx.x & 31u;
    if (tile == 0) return 0;
    uint64_t excl = 0;
    int64_t base = (int64_t)tile - 1;           // lane l looks at tile base - l
    for (;;) {
        const int64_t t = base - (int64_t)lane;
        uint64_t v = t >= 0 ? ld_volatile_u64(state + t) : SIGK_CS_PRE;     // before tile 0: prefix 0
        // the nearest inclusive prefix ends the walk; everything nearer must be published first
        unsigned pre = __ballot_sync(0xffffffffu, (v >> 62) == 2);
        unsigned none = __ballot_sync(0xffffffffu, (v >> 62) == 0);
        const unsigned upto = pre ? (pre & (0u - pre)) : 0u;               // lowest lane holding a prefix
        const unsigned needed = upto ? ((upto << 1) - 1u) : 0xffffffffu;    // lanes 0 .. that lane
        if (none & needed) { __nanosleep(40); continue; }                   // somebody nearer is not ready: reread
        uint64_t c = ((1u << lane) & needed) ? (v & SIGK_CS_VAL) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        excl += c;
        if (pre) break;
        base -= 32;
    }
    if (lane == 0) st_volatile_u64(state + tile, SIGK_CS_PRE | (excl + aggregate));
    return excl;
}
SIGK_D uint64_t chained_scan_exclusive_warp(uint64_t *state, uint32_t tile, uint64_t aggregate) {
    chained_scan_publish_warp(state, tile, aggregate);
    return chained_scan_resolve_warp(state, tile, aggregate);
}

// Block-wide exclusive scan of one uint32 per thread; returns the exclusive
// prefix, writes the block total to *total (valid in every thread).
template <int THREADS>
SIGK_D uint32_t block_exclusive_scan(uint32_t x, uint32_t *s_warp /* THREADS/32 + 1 words */, uint32_t *total) {
    constexpr int WARPS = THREADS / 32;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < WARPS ? s_warp[lane] : 0;
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= (unsigned)o) wi += y;
        }
        if (lane < WARPS) s_warp[lane] = wi - w;
        if (lane == WARPS - 1) s_warp[WARPS] = wi;
    }
    __syncthreads();
    const uint32_t res = s_warp[warp] + incl - x;
    *total = s_warp[WARPS];
    __syncthreads();   // s_warp may be reused by the caller
    return res;
}

#endif  // __CUDACC__
