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

enum : uint32_t { S_START = 0, S_ID = 1, S_DEF = 2, S_DATA = 3, S_LINE = 4 };
enum : uint32_t { C_CR = 0, C_NL = 1, C_GT = 2, C_BLANK = 3, C_ALPHA = 4, C_STAR = 5, C_OTHER = 6 };

constexpr uint32_t pack5(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e) { return a | b << 3 | c << 6 | d << 9 | e << 12; }
constexpr uint32_t FN_IDENTITY = pack5(S_START, S_ID, S_DEF, S_DATA, S_LINE);
// next state per class, indexed by the current state (START, ID, DEF, DATA, LINE)
__constant__ uint32_t c_fa_next[7] = {
    FN_IDENTITY,                                        // '\r'
    pack5(S_START, S_DATA, S_DATA, S_LINE, S_LINE),     // '\n'
    pack5(S_ID, S_ID, S_DEF, S_DATA, S_ID),             // '>'
    pack5(S_START, S_DEF, S_DEF, S_DATA, S_LINE),       // blank
    pack5(S_START, S_ID, S_DEF, S_DATA, S_DATA),        // letter
    FN_IDENTITY,                                        // '*'
    FN_IDENTITY,                                        // anything else
};

SIGK_D uint32_t fa_class(uint32_t c) {
    if (c == '\r') return C_CR;
    if (c == '\n') return C_NL;
    if (c == '>') return C_GT;
    if (c == ' ' || c == '\t') return C_BLANK;
    if (((c | 0x20u) - 'a') < 26u) return C_ALPHA;      // isalpha in the C locale
    if (c == '*') return C_STAR;
    return C_OTHER;
}
SIGK_D uint32_t fa_apply(uint32_t f, uint32_t s) { return (f >> (3u * s)) & 7u; }
// first f, then g
SIGK_D uint32_t fa_compose(uint32_t f, uint32_t g) {
    uint32_t r = 0;
#pragma unroll
    for (uint32_t s = 0; s < 5; ++s) r |= fa_apply(g, fa_apply(f, s)) << (3u * s);
    return r;
}
constexpr uint32_t FN_ONES = pack5(1, 1, 1, 1, 1);
SIGK_D uint32_t fa_constant(uint32_t s) { return s * FN_ONES; }

// what a byte of class c does when the machine is in state s
struct FaAction { bool seq, record, id_end, line_end, error; };
SIGK_D FaAction fa_action(uint32_t s, uint32_t c) {
    FaAction a;
    a.seq = (s == S_DATA && (c == C_ALPHA || c == C_STAR)) || (s == S_LINE && c == C_ALPHA);
    a.record = (s == S_START || s == S_LINE) && c == C_GT;
    a.id_end = s == S_ID && (c == C_BLANK || c == C_NL);
    a.line_end = (s == S_ID || s == S_DEF) && c == C_NL;
    a.error = c != C_CR && ((s == S_START && c != C_GT) || (s == S_DATA && (c == C_BLANK || c == C_GT || c == C_OTHER)) ||
                            (s == S_LINE && (c == C_STAR || c == C_BLANK || c == C_OTHER)));
    return a;
}

// the thread's 16 bytes (bytes past the tile's end read as '\r': no transition, no output)
SIGK_D void fa_load(const uint8_t *__restrict__ bytes, const FastaTile &t, uint32_t cls[FA_BPT], uint8_t raw[FA_BPT]) {
    const uint32_t off = threadIdx.x * FA_BPT;
    uint4 q = make_uint4(0, 0, 0, 0);
    if (off < t.n) q = *reinterpret_cast<const uint4 *>(bytes + t.begin + off);      // (file starts are 16-byte aligned, the buffer is padded)
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < FA_BPT; ++i) {
        const uint32_t c = (w[i >> 2] >> (8 * (i & 3))) & 0xFFu;
        raw[i] = (uint8_t)c;
        cls[i] = off + i < t.n ? fa_class(c) : (uint32_t)C_CR;
    }
}
SIGK_D uint32_t fa_thread_fn(const uint32_t cls[FA_BPT]) {
    uint32_t f = FN_IDENTITY;
#pragma unroll
    for (int i = 0; i < FA_BPT; ++i) f = fa_compose(f, c_fa_next[cls[i]]);
    return f;
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
    if (total) {
        uint32_t all = before;
        for (unsigned w = warp; w < FA_WARPS; ++w) all = fa_compose(all, s_warp[w]);
        *total = all;
    }
    __syncthreads();
    return fa_compose(before, excl);
}

// ---- sweep 1: the transition function of every tile -------------------------------------------------------------
__global__ void __launch_bounds__(FA_THREADS)
fasta_tile_fn_kernel(const uint8_t *__restrict__ bytes, const FastaTile *__restrict__ tiles, uint32_t *__restrict__ tile_fn) {
    __shared__ uint32_t s_warp[FA_WARPS];
    const FastaTile t = tiles[blockIdx.x];
    uint32_t cls[FA_BPT];
    uint8_t raw[FA_BPT];
    fa_load(bytes, t, cls, raw);
    uint32_t total;
    fa_block_scan_fn(fa_thread_fn(cls), s_warp, &total);
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
    __syncthreads();
    return before + incl - v;
}

// ---- sweep 2 (COUNT) and 3 (EMIT): every byte with its entry state ------------------------------------------------
template <bool EMIT>
__global__ void __launch_bounds__(FA_THREADS)
fasta_sweep_kernel(const uint8_t *__restrict__ bytes, const FastaTile *__restrict__ tiles, const uint8_t *__restrict__ tile_state,
                   uint64_t *__restrict__ tile_counts,         // COUNT: out, packed; EMIT: in, exclusive prefix {seq, records, errors} per tile
                   FastaOut out) {
    __shared__ uint32_t s_warp[FA_WARPS];
    __shared__ uint64_t s_warp64[FA_WARPS];
    const FastaTile t = tiles[blockIdx.x];
    uint32_t cls[FA_BPT];
    uint8_t raw[FA_BPT];
    fa_load(bytes, t, cls, raw);
    const uint32_t before = fa_block_scan_fn(fa_thread_fn(cls), s_warp, nullptr);
    const uint32_t s0 = fa_apply(before, tile_state[blockIdx.x]);
    uint32_t s = s0, n_seq = 0, n_rec = 0, n_err = 0;
#pragma unroll
    for (int i = 0; i < FA_BPT; ++i) {
        const FaAction a = fa_action(s, cls[i]);
        n_seq += a.seq; n_rec += a.record; n_err += a.error;
        s = fa_apply(c_fa_next[cls[i]], s);
    }
    uint64_t total;
    const uint64_t excl = fa_block_scan_u64((uint64_t)n_seq | (uint64_t)n_rec << 16 | (uint64_t)n_err << 32, s_warp64, &total);
    if (!EMIT) {
        if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
        return;
    }
    uint64_t seq_at = tile_counts[3 * (size_t)blockIdx.x] + (excl & 0xFFFFu);
    uint64_t rec_at = tile_counts[3 * (size_t)blockIdx.x + 1] + ((excl >> 16) & 0xFFFFu);      // records opened before this byte
    uint64_t err_at = tile_counts[3 * (size_t)blockIdx.x + 2] + (excl >> 32);
    const uint64_t pos0 = t.begin + (uint64_t)threadIdx.x * FA_BPT;
    s = s0;
#pragma unroll
    for (int i = 0; i < FA_BPT; ++i) {
        const FaAction a = fa_action(s, cls[i]);
        const uint64_t pos = pos0 + i;
        if (a.record) {
            out.header_pos[rec_at] = pos;
            out.seq_begin[rec_at] = seq_at;
            ++rec_at;
        }
        if (a.seq) out.residues[seq_at++] = raw[i];
        if (a.id_end) out.id_end[rec_at - 1] = pos;            // (a byte seen in s_id or s_defline follows its record's '>')
        if (a.line_end) out.line_end[rec_at - 1] = pos;
        if (a.error) {
            if (err_at < out.err_capacity) { out.err_pos[err_at] = pos | (uint64_t)s << 60; out.err_record[err_at] = rec_at ? (uint32_t)(rec_at - 1) : 0xFFFFFFFFu; }
            ++err_at;
        }
        s = fa_apply(c_fa_next[cls[i]], s);
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
    for (uint64_t i = lane; i < n; i += 32) residues[dst + i] = stream[src + i];
}

}  // namespace

cudaError_t launch_fasta_tile_functions(const uint8_t *bytes, const FastaTile *tiles, uint32_t n_tiles, uint32_t *tile_fn, uint8_t *tile_state,
                                        cudaStream_t stream) {
    if (n_tiles == 0) return cudaSuccess;
    fasta_tile_fn_kernel<<<n_tiles, FA_THREADS, 0, stream>>>(bytes, tiles, tile_fn);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    fasta_state_scan_kernel<<<1, FS_THREADS, 0, stream>>>(tile_fn, tiles, n_tiles, tile_state);
    return cudaGetLastError();
}

cudaError_t launch_fasta_count(const uint8_t *bytes, const FastaTile *tiles, uint32_t n_tiles, const uint8_t *tile_state, uint64_t *tile_packed,
                               uint64_t *tile_prefix, uint64_t *totals, cudaStream_t stream) {
    if (n_tiles == 0) return cudaMemsetAsync(totals, 0, 3 * sizeof(uint64_t), stream);
    fasta_sweep_kernel<false><<<n_tiles, FA_THREADS, 0, stream>>>(bytes, tiles, tile_state, tile_packed, FastaOut{});
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    fasta_count_scan_kernel<<<1, FS_THREADS, 0, stream>>>(tile_packed, n_tiles, tile_prefix, totals);
    return cudaGetLastError();
}

cudaError_t launch_fasta_emit(const uint8_t *bytes, const FastaTile *tiles, uint32_t n_tiles, const uint8_t *tile_state, uint64_t *tile_prefix,
                              const FastaOut &out, cudaStream_t stream) {
    if (n_tiles == 0) return cudaSuccess;
    fasta_sweep_kernel<true><<<n_tiles, FA_THREADS, 0, stream>>>(bytes, tiles, tile_state, tile_prefix, out);
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
