// fasta.cu — FASTA bytes -> record table + packed residues, on the device.
//
// Replaces the character-at-a-time state machine of the reference's FastaParser::parse_char
// (src/fasta_parser.h:38-144) as it is driven by SignatureBuilder<K>::load_kmers_from_fasta
// (src/signature_build.tcc:84-102): five states, every quirk kept —
//   '\r' is dropped before the state is looked at (:46-47);
//   s_start: anything but '>' is an error and the state stays (:52-61);
//   s_id: a blank (isblank: ' ' or '\t') ends the id and is the first character of the definition, '\n' ends
//         the header, everything else — '>' included — is an id character (:63-77);
//   s_defline: everything up to '\n' (:79-88);
//   s_data: '\n' -> s_id_or_data; letters and '*' are sequence; anything else is reported and dropped, '>' included,
//           so the header that follows a header-only record lands in that record's sequence (:90-106);
//   s_id_or_data: '>' closes the record and opens the next, '\n' stays, a letter is sequence, anything else
//           ('*' included) is reported and dropped (:108-133).
// A state machine over bytes is a composition of per-byte transition functions, and composition is
// associative: a function here is five 3-bit states packed in 15 bits, a tile composes its bytes'
// functions, one block scans the tiles' functions (files restart in s_start), and a second sweep over
// the bytes — now with the entry state of every byte known — classifies each byte and scatters
//   the sequence characters of all records, packed, in file order (the residue stream);
//   per record: where its '>' is, where its id ends, where its header line ends, where its
//   sequence starts in the residue stream;
//   the positions (and states) of the reported characters.
// The caller (host) turns ids into function indices — a string-keyed map, FunctionMap::lookup_function —
// and sigk_fasta_commit gathers the records it keeps into the build's input arrays without the
// residues ever crossing PCIe again.
#include "kernels.h"
#include "sigk_common.cuh"

namespace sigk {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int FA_THREADS = 512;
constexpr int FA_BPT = 16;                          // bytes per thread: one 128-bit load
constexpr int FA_WARPS = FA_THREADS / 32;
static_assert(FA_THREADS * FA_BPT == FASTA_TILE, "tile size");

// A state is kept as its shift amount 5 * s, so that "look the state up in a packed table" is a shift and a mask:
// a transition function is five 5-bit fields (the image of state s in bits [5s, 5s+5)), and so is the table of
// what a byte class does in every state.
enum : uint32_t { S_START = 0, S_ID = 5, S_DEF = 10, S_DATA = 15, S_LINE = 20 };
enum : uint32_t { C_CR = 0, C_NL = 1, C_GT = 2, C_BLANK = 3, C_ALPHA = 4, C_STAR = 5, C_OTHER = 6 };
enum : uint32_t { A_SEQ = 1, A_RECORD = 2, A_ID_END = 4, A_LINE_END = 8, A_ERROR = 16 };

constexpr uint32_t pack5(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e) { return a | b << 5 | c << 10 | d << 15 | e << 20; }
constexpr uint32_t FN_IDENTITY = pack5(S_START, S_ID, S_DEF, S_DATA, S_LINE);
constexpr uint32_t FN_ONES = pack5(1, 1, 1, 1, 1);
// per class: the next state and what the byte does, indexed by the current state (START, ID, DEF, DATA, LINE)
constexpr uint32_t FA_NEXT[7] = {
    FN_IDENTITY,                                        // '\r'
    pack5(S_START, S_DATA, S_DATA, S_LINE, S_LINE),     // '\n'
    pack5(S_ID, S_ID, S_DEF, S_DATA, S_ID),             // '>'
    pack5(S_START, S_DEF, S_DEF, S_DATA, S_LINE),       // blank
    pack5(S_START, S_ID, S_DEF, S_DATA, S_DATA),        // letter
    FN_IDENTITY,                                        // '*'
    FN_IDENTITY,                                        // anything else
};
constexpr uint32_t FA_ACT[7] = {
    0,                                                                  // '\r': dropped before the state is looked at
    pack5(A_ERROR, A_ID_END | A_LINE_END, A_LINE_END, 0, 0),            // '\n'
    pack5(A_RECORD, 0, 0, A_ERROR, A_RECORD),                           // '>'
    pack5(A_ERROR, A_ID_END, 0, A_ERROR, A_ERROR),                      // blank
    pack5(A_ERROR, 0, 0, A_SEQ, A_SEQ),                                 // letter
    pack5(A_ERROR, 0, 0, A_SEQ, A_ERROR),                               // '*'
    pack5(A_ERROR, 0, 0, A_ERROR, A_ERROR),                             // anything else
};
__constant__ uint32_t c_fa_next[7] = {FA_NEXT[0], FA_NEXT[1], FA_NEXT[2], FA_NEXT[3], FA_NEXT[4], FA_NEXT[5], FA_NEXT[6]};
__constant__ uint32_t c_fa_act[7] = {FA_ACT[0], FA_ACT[1], FA_ACT[2], FA_ACT[3], FA_ACT[4], FA_ACT[5], FA_ACT[6]};

SIGK_D uint32_t fa_class(uint32_t c) {
    if (c == '\r') return C_CR;
    if (c == '\n') return C_NL;
    if (c == '>') return C_GT;
    if (c == ' ' || c == '\t') return C_BLANK;
    if (((c | 0x20u) - 'a') < 26u) return C_ALPHA;      // isalpha in the C locale
    if (c == '*') return C_STAR;
    return C_OTHER;
}
SIGK_D uint32_t fa_apply(uint32_t f, uint32_t s) { return (f >> s) & 31u; }
// first f, then g
SIGK_D uint32_t fa_compose(uint32_t f, uint32_t g) {
    uint32_t r = 0;
#pragma unroll
    for (uint32_t i = 0; i < 25; i += 5) r |= fa_apply(g, fa_apply(f, i)) << i;
    return r;
}
SIGK_D uint32_t fa_constant(uint32_t s) { return s * FN_ONES; }

// byte -> {next-state table, action table} in shared memory: one 64-bit load per byte, off the state chain
SIGK_D void fa_fill_lut(uint2 *s_lut) {
    if (threadIdx.x < 256) {
        const uint32_t c = fa_class(threadIdx.x);
        s_lut[threadIdx.x] = make_uint2(c_fa_next[c], c_fa_act[c]);
    }
    __syncthreads();
}
// the thread's 16 bytes (bytes past the tile's end behave like '\r': no transition, no output)
SIGK_D void fa_load(const uint8_t *__restrict__ bytes, const FastaTile &t, const uint2 *s_lut, uint32_t nx[FA_BPT], uint32_t act[FA_BPT], uint4 &q) {
    const uint32_t off = threadIdx.x * FA_BPT;
    q = make_uint4(0, 0, 0, 0);
    if (off < t.n) q = *reinterpret_cast<const uint4 *>(bytes + t.begin + off);      // (file starts are 16-byte aligned, the buffer is padded)
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < FA_BPT; ++i) {
        const uint2 e = s_lut[(w[i >> 2] >> (8 * (i & 3))) & 0xFFu];
        const bool in = off + i < t.n;
        nx[i] = in ? e.x : FN_IDENTITY;
        act[i] = in ? e.y : 0u;
    }
}
SIGK_D uint32_t fa_byte(const uint4 &q, int i) {
    const uint32_t w = i < 4 ? q.x : i < 8 ? q.y : i < 12 ? q.z : q.w;
    return (w >> (8 * (i & 3))) & 0xFFu;
}

// exclusive scan of the threads' functions in thread order; *total = the tile's function
SIGK_D uint32_t fa_block_scan_fn(uint32_t f, uint32_t *s_warp, uint32_t *total) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t incl = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(FULL, incl, o);
        if (lane >= (unsigned)o) incl = fa_compose(up, incl);
    }
    uint32_t excl = __shfl_up_sync(FULL, incl, 1);
    if (lane == 0) excl = FN_IDENTITY;
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t before = FN_IDENTITY;
    for (unsigned w = 0; w < warp; ++w) before = fa_compose(before, s_warp[w]);
    uint32_t all = before;
    for (unsigned w = warp; w < FA_WARPS; ++w) all = fa_compose(all, s_warp[w]);
    *total = all;
    return fa_compose(before, excl);
}

// ---- sweep 1: every 16-byte chunk's function, scanned inside its tile; the transition function of every tile -----
__global__ void __launch_bounds__(FA_THREADS)
fasta_tile_fn_kernel(const uint8_t *__restrict__ bytes, const FastaTile *__restrict__ tiles, uint32_t *__restrict__ chunk_before,
                     uint32_t *__restrict__ tile_fn) {
    __shared__ uint32_t s_warp[FA_WARPS];
    __shared__ uint2 s_lut[256];
    fa_fill_lut(s_lut);
    const FastaTile t = tiles[blockIdx.x];
    uint32_t nx[FA_BPT], act[FA_BPT];
    uint4 q;
    fa_load(bytes, t, s_lut, nx, act, q);
    // the chunk's function: where each of the five states ends up, five independent chains
    uint32_t s0 = S_START, s1 = S_ID, s2 = S_DEF, s3 = S_DATA, s4 = S_LINE;
#pragma unroll
    for (int i = 0; i < FA_BPT; ++i) {
        s0 = fa_apply(nx[i], s0); s1 = fa_apply(nx[i], s1); s2 = fa_apply(nx[i], s2); s3 = fa_apply(nx[i], s3); s4 = fa_apply(nx[i], s4);
    }
    uint32_t total;
    const uint32_t before = fa_block_scan_fn(s0 | s1 << 5 | s2 << 10 | s3 << 15 | s4 << 20, s_warp, &total);
    chunk_before[(size_t)blockIdx.x * FA_THREADS + threadIdx.x] = before;       // what the tile's earlier chunks do, composed
    if (threadIdx.x == 0) tile_fn[blockIdx.x] = total;
}

// ---- the state every tile starts in: one block, every thread a run of consecutive tiles ---------------------------
constexpr int FS_THREADS = 1024;
__global__ void __launch_bounds__(FS_THREADS)
fasta_state_scan_kernel(const uint32_t *__restrict__ tile_fn, const FastaTile *__restrict__ tiles, uint32_t n_tiles, uint8_t *__restrict__ tile_state) {
    __shared__ uint32_t s_fn[FS_THREADS];
    const uint32_t per = (n_tiles + FS_THREADS - 1) / FS_THREADS;
    const uint32_t lo = min(n_tiles, threadIdx.x * per), hi = min(n_tiles, lo + per);
    // a file's first tile starts in s_start whatever came before: its function, seen from the previous tiles, is constant
    uint32_t f = FN_IDENTITY;
    for (uint32_t t = lo; t < hi; ++t) f = tiles[t].file_start ? fa_constant(fa_apply(tile_fn[t], S_START)) : fa_compose(f, tile_fn[t]);
    s_fn[threadIdx.x] = f;
    __syncthreads();
    if (threadIdx.x == 0) {                     // 1024 compositions, once per parse
        uint32_t run = FN_IDENTITY;
        for (int i = 0; i < FS_THREADS; ++i) { const uint32_t mine = s_fn[i]; s_fn[i] = run; run = fa_compose(run, mine); }
    }
    __syncthreads();
    uint32_t s = fa_apply(s_fn[threadIdx.x], S_START);
    for (uint32_t t = lo; t < hi; ++t) {
        if (tiles[t].file_start) s = S_START;
        tile_state[t] = (uint8_t)s;
        s = fa_apply(tile_fn[t], s);
    }
}

// packed per-thread counts for one scan: sequence characters | records << 16 | reported characters << 32
SIGK_D uint64_t fa_block_scan_u64(uint64_t v, uint64_t *s_warp, uint64_t *total) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t up = __shfl_up_sync(FULL, incl, o);
        if (lane >= (unsigned)o) incl += up;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint64_t before = 0, all = 0;
    for (unsigned w = 0; w < FA_WARPS; ++w) { if (w < warp) before += s_warp[w]; all += s_warp[w]; }
    *total = all;
    return before + incl - v;
}

// ---- sweep 2 (COUNT) and 3 (EMIT): every byte with its entry state ------------------------------------------------
// COUNT leaves every chunk's counts (5 + 4 + 5 bits) for EMIT, which scans them and walks its bytes once.
template <bool EMIT>
__global__ void __launch_bounds__(FA_THREADS)
fasta_sweep_kernel(const uint8_t *__restrict__ bytes, const FastaTile *__restrict__ tiles, const uint8_t *__restrict__ tile_state,
                   const uint32_t *__restrict__ chunk_before, uint16_t *__restrict__ chunk_counts,
                   uint64_t *__restrict__ tile_counts,         // COUNT: out, packed; EMIT: in, exclusive prefix {seq, records, errors} per tile
                   FastaOut out) {
    __shared__ uint64_t s_warp64[FA_WARPS];
    __shared__ uint2 s_lut[256];
    fa_fill_lut(s_lut);
    const FastaTile t = tiles[blockIdx.x];
    const size_t chunk = (size_t)blockIdx.x * FA_THREADS + threadIdx.x;
    uint32_t nx[FA_BPT], act[FA_BPT];
    uint4 q;
    fa_load(bytes, t, s_lut, nx, act, q);
    uint32_t s = fa_apply(chunk_before[chunk], tile_state[blockIdx.x]);
    if (!EMIT) {
        uint32_t n_seq = 0, n_rec = 0, n_err = 0;
#pragma unroll
        for (int i = 0; i < FA_BPT; ++i) {
            const uint32_t a = fa_apply(act[i], s);
            n_seq += a & 1u; n_rec += (a >> 1) & 1u; n_err += a >> 4;
            s = fa_apply(nx[i], s);
        }
        chunk_counts[chunk] = (uint16_t)(n_seq | n_rec << 5 | n_err << 9);
        uint64_t total;
        fa_block_scan_u64((uint64_t)n_seq | (uint64_t)n_rec << 16 | (uint64_t)n_err << 32, s_warp64, &total);
        if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
        return;
    }
    const uint32_t cc = chunk_counts[chunk];
    uint64_t total;
    const uint64_t excl = fa_block_scan_u64((uint64_t)(cc & 31u) | (uint64_t)((cc >> 5) & 15u) << 16 | (uint64_t)(cc >> 9) << 32, s_warp64, &total);
    uint64_t seq_at = tile_counts[3 * (size_t)blockIdx.x] + (excl & 0xFFFFu);
    uint64_t rec_at = tile_counts[3 * (size_t)blockIdx.x + 1] + ((excl >> 16) & 0xFFFFu);      // records opened before this byte
    uint64_t err_at = tile_counts[3 * (size_t)blockIdx.x + 2] + (excl >> 32);
    const uint64_t pos0 = t.begin + (uint64_t)threadIdx.x * FA_BPT;
#pragma unroll
    for (int i = 0; i < FA_BPT; ++i) {
        const uint32_t a = fa_apply(act[i], s);
        const uint64_t pos = pos0 + i;
        if (a & A_RECORD) {
            out.header_pos[rec_at] = pos;
            out.seq_begin[rec_at] = seq_at;
            ++rec_at;
        }
        if (a & A_SEQ) out.residues[seq_at++] = (uint8_t)fa_byte(q, i);
        if (a & A_ID_END) out.id_end[rec_at - 1] = pos;        // (a byte seen in s_id or s_defline follows its record's '>')
        if (a & A_LINE_END) out.line_end[rec_at - 1] = pos;
        if (a & A_ERROR) {
            if (err_at < out.err_capacity) { out.err_pos[err_at] = pos | (uint64_t)(s / 5u) << 60; out.err_record[err_at] = rec_at ? (uint32_t)(rec_at - 1) : 0xFFFFFFFFu; }
            ++err_at;
        }
        s = fa_apply(nx[i], s);
    }
}

// ---- exclusive scan of the tiles' packed counts into three 64-bit prefixes per tile; totals[3] ---------------------
__global__ void __launch_bounds__(FS_THREADS)
fasta_count_scan_kernel(const uint64_t *__restrict__ packed, uint32_t n_tiles, uint64_t *__restrict__ prefix, uint64_t *__restrict__ totals) {
    __shared__ uint64_t s_sum[3][FS_THREADS];
    const uint32_t per = (n_tiles + FS_THREADS - 1) / FS_THREADS;
    const uint32_t lo = min(n_tiles, threadIdx.x * per), hi = min(n_tiles, lo + per);
    uint64_t a = 0, b = 0, c = 0;
    for (uint32_t t = lo; t < hi; ++t) { const uint64_t v = packed[t]; a += v & 0xFFFFu; b += (v >> 16) & 0xFFFFu; c += v >> 32; }
    s_sum[0][threadIdx.x] = a; s_sum[1][threadIdx.x] = b; s_sum[2][threadIdx.x] = c;
    __syncthreads();
    if (threadIdx.x < 3) {
        uint64_t run = 0;
        for (int i = 0; i < FS_THREADS; ++i) { const uint64_t mine = s_sum[threadIdx.x][i]; s_sum[threadIdx.x][i] = run; run += mine; }
        totals[threadIdx.x] = run;
    }
    __syncthreads();
    a = s_sum[0][threadIdx.x]; b = s_sum[1][threadIdx.x]; c = s_sum[2][threadIdx.x];
    for (uint32_t t = lo; t < hi; ++t) {
        const uint64_t v = packed[t];
        prefix[3 * (size_t)t] = a; prefix[3 * (size_t)t + 1] = b; prefix[3 * (size_t)t + 2] = c;
        a += v & 0xFFFFu; b += (v >> 16) & 0xFFFFu; c += v >> 32;
    }
}

// ---- commit: the kept records' residues, gathered into the build's residue array in record order -------------------
// One warp per protein: src = its record's slice of the residue stream, dst = its slice of the packed input.
__global__ void __launch_bounds__(256)
fasta_gather_kernel(const uint8_t *__restrict__ stream, const uint64_t *__restrict__ src_begin, const uint64_t *__restrict__ starts,
                    uint32_t n_proteins, uint8_t *__restrict__ residues) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (warp >= n_proteins) return;
    const uint64_t dst = starts[warp], n = starts[warp + 1] - dst, src = src_begin[warp];
    uint64_t i = lane;
    for (; i + 96 < n; i += 128) {             // four loads in flight per lane
        const uint8_t a = stream[src + i], b = stream[src + i + 32], c = stream[src + i + 64], d = stream[src + i + 96];
        residues[dst + i] = a; residues[dst + i + 32] = b; residues[dst + i + 64] = c; residues[dst + i + 96] = d;
    }
    for (; i < n; i += 32) residues[dst + i] = stream[src + i];
}

}  // namespace

cudaError_t launch_fasta_tile_functions(const uint8_t *bytes, const FastaTile *tiles, uint32_t n_tiles, uint32_t *chunk_before, uint32_t *tile_fn,
                                        uint8_t *tile_state, cudaStream_t stream) {
    if (n_tiles == 0) return cudaSuccess;
    fasta_tile_fn_kernel<<<n_tiles, FA_THREADS, 0, stream>>>(bytes, tiles, chunk_before, tile_fn);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    fasta_state_scan_kernel<<<1, FS_THREADS, 0, stream>>>(tile_fn, tiles, n_tiles, tile_state);
    return cudaGetLastError();
}

cudaError_t launch_fasta_count(const uint8_t *bytes, const FastaTile *tiles, uint32_t n_tiles, const uint8_t *tile_state, const uint32_t *chunk_before,
                               uint16_t *chunk_counts, uint64_t *tile_packed, uint64_t *tile_prefix, uint64_t *totals, cudaStream_t stream) {
    if (n_tiles == 0) return cudaMemsetAsync(totals, 0, 3 * sizeof(uint64_t), stream);
    fasta_sweep_kernel<false><<<n_tiles, FA_THREADS, 0, stream>>>(bytes, tiles, tile_state, chunk_before, chunk_counts, tile_packed, FastaOut{});
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    fasta_count_scan_kernel<<<1, FS_THREADS, 0, stream>>>(tile_packed, n_tiles, tile_prefix, totals);
    return cudaGetLastError();
}

cudaError_t launch_fasta_emit(const uint8_t *bytes, const FastaTile *tiles, uint32_t n_tiles, const uint8_t *tile_state, const uint32_t *chunk_before,
                              uint16_t *chunk_counts, uint64_t *tile_prefix, const FastaOut &out, cudaStream_t stream) {
    if (n_tiles == 0) return cudaSuccess;
    fasta_sweep_kernel<true><<<n_tiles, FA_THREADS, 0, stream>>>(bytes, tiles, tile_state, chunk_before, chunk_counts, tile_prefix, out);
    return cudaGetLastError();
}

cudaError_t launch_fasta_gather(const uint8_t *stream_bytes, const uint64_t *src_begin, const uint64_t *starts, uint32_t n_proteins,
                                uint8_t *residues, cudaStream_t stream) {
    if (n_proteins == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)(((uint64_t)n_proteins * 32 + 255) / 256);
    fasta_gather_kernel<<<blocks, 256, 0, stream>>>(stream_bytes, src_begin, starts, n_proteins, residues);
    return cudaGetLastError();
}

}  // namespace sigk
