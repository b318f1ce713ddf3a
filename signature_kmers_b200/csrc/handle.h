// handle.h — the libsigk handle (private to csrc/): device buffers, stream,
// events, and the multi-GPU communicator state.
#pragma once

#include "../../include/sigk.h"
#include "kernels.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

namespace sigk {

struct DeviceScalars {
    uint64_t n_records;         // valid windows this GPU sorts: n_main + n_side
    uint64_t n_segments;
    uint64_t n_kept;
    uint64_t n_seqs_sig;
    uint64_t n_main, n_side, n_both;   // written together by scan_bins_kernel: records without / with a lower-case residue
    uint64_t n_side_kept;       // kept rows of the side run (the table's second section)
    uint32_t overflow, pad0;
    uint32_t ticket[16];
    uint32_t n_groups, next_group, n_long, next_long, n_work, next_work, n_work_long, next_work_long, n_giant, next_giant;
    uint64_t reduce_in[4];      // multi-GPU: {occurrences, groups, kept, -} of this rank -> summed over ranks
};

enum { TK_ENCODE = 0, TK_REDUCE = 1, TK_SQUEEZE = 2, TK_SORT0 = 3, TK_SIDE0 = 8 };      // TK_SORT0 + main passes <= TK_SIDE0; TK_SIDE0 + side passes <= 16

template <typename T> struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;     // elements
    cudaError_t reserve(size_t n) {
        if (p && n <= cap) return cudaSuccess;      // reserve(0) still yields a valid pointer (empty ranks, empty inputs)
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

template <typename T> struct PinnedBuf {
    T *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (p && n <= cap) return cudaSuccess;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        cudaError_t e = cudaMallocHost(&p, std::max<size_t>(n, 1) * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

enum { EV_START, EV_H2D, EV_DEV0, EV_ENCODE, EV_EXCHANGE, EV_HIST, EV_MAIN_SORTED, EV_SORT, EV_RED_COUNT, EV_RED_EMIT, EV_REJ0, EV_REJ1, EV_REDUCE, EV_ORDER, EV_SQUEEZE, EV_D2H,
       EV_USER0, EV_USER1, EV_USER2, EV_USER3, EV_FA0, EV_FA1, EV_FA2, EV_FA3, EV_PASS0, EV_COUNT = EV_PASS0 + SORT_MAX_PASSES + 1 };

struct Comm;    // comm.cu

}  // namespace sigk

struct sigk_handle {
    sigk_config cfg{};
    std::string error;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[sigk::EV_COUNT] = {};
    int sm_count = 148;

    sigk_proteins in{};
    bool have_input = false, uploaded = false, built = false, downloaded = false;
    uint64_t total_res = 0;
    uint32_t max_seq_id = 0;            // over the whole job once a communicator is joined
    uint32_t local_max_seq_id = 0;
    uint32_t local_max_function = 0;    // largest function_index among this rank's proteins

    sigk::DevBuf<uint8_t> d_res;
    sigk::DevBuf<uint64_t> d_starts;
    sigk::DevBuf<uint16_t> d_func;
    sigk::DevBuf<uint32_t> d_seqid, d_slice_prot;
    sigk::DevBuf<uint8_t> d_meta;          // per-protein {length, function}: 4 or 8 bytes each (meta_compact)
    bool meta_compact = false;
    int rej_shift = 0;                     // SIGK_TEST_REJ_SPREAD: the same for the per-protein rejected-occurrence counters (single GPU only)
    int meta_shift = 0;                    // SIGK_TEST_META_SPREAD: table entries 2^shift apart (cache-footprint experiments)
    uint64_t local_max_len = 0, max_len = 0;   // longest protein: this rank's / the job's
    sigk::DevBuf<uint4> d_rows;
    sigk::DevBuf<sigk::OrderWork> d_groups, d_long_groups, d_work, d_work_long, d_giant;
    sigk::DevBuf<uint64_t> d_keys[2];
    sigk::DevBuf<uint32_t> d_vals[2];
    sigk::DevBuf<uint8_t> d_lookback;
    sigk::DevBuf<uint64_t> d_hist, d_binbase, d_scan_state;
    sigk::DevBuf<uint64_t> d_out_kmer;
    // sigk_lookup: query batch and its result, grow-only; table_on_device = a table (built or set) is resident
    sigk::DevBuf<uint8_t> d_q_res;
    sigk::DevBuf<uint64_t> d_q_starts;
    sigk::DevBuf<uint32_t> d_q_rows;
    bool table_on_device = false;
    sigk::DevBuf<uint16_t> d_out_cols;       // 5 columns of capacity rows
    sigk::DevBuf<uint32_t> d_bitmap, d_distinct, d_swf, d_prot_windows, d_prot_rejected;
    sigk::DevBuf<sigk::DeviceScalars> d_scalars;
    uint64_t capacity = 0;             // records the buffers are sized for
    int sorted_in = 0;                 // which ping-pong buffer holds the sorted records

    sigk::PinnedBuf<uint64_t> h_kmer;
    sigk::PinnedBuf<uint16_t> h_cols;
    sigk::PinnedBuf<uint32_t> h_distinct, h_swf;
    sigk::PinnedBuf<sigk::DeviceScalars> h_scalars;
    uint64_t h_rows = 0;

    // sigk_fasta_parse / sigk_fasta_commit (fasta.cu)
    bool input_on_device = false;      // the input arrays were packed on the device by sigk_fasta_commit: sigk_upload copies nothing
    bool fasta_parsed = false;
    sigk::DevBuf<uint8_t> d_fa_bytes, d_fa_state, d_fa_stream;
    sigk::DevBuf<sigk::FastaTile> d_fa_tiles;
    sigk::DevBuf<uint32_t> d_fa_fn, d_fa_err_rec, d_fa_chunk_fn;
    sigk::DevBuf<uint16_t> d_fa_chunk_counts;
    sigk::DevBuf<uint64_t> d_fa_packed, d_fa_prefix, d_fa_totals, d_fa_rec, d_fa_err_pos, d_fa_src;     // d_fa_rec: 4 arrays of (records + 1)
    sigk::PinnedBuf<uint64_t> h_fa_rec, h_fa_totals, h_fa_err_pos, h_fa_starts, h_fa_src;
    sigk::PinnedBuf<uint16_t> h_fa_func;
    sigk::PinnedBuf<uint32_t> h_fa_sid;
    sigk::PinnedBuf<uint32_t> h_fa_err_rec;
    uint64_t fa_records = 0, fa_residues = 0, fa_errors = 0, fa_rec_stride = 0, fa_bytes = 0;

    sigk_timings tm{};
    float h2d_ms = 0;
    sigk::PassPlan plan{}, plan_side{};
    bool fused = true;                 // single GPU: encode fused with the first radix pass (SIGK_NO_FUSED=1 turns it off)

    // multi-GPU
    sigk::Comm *comm = nullptr;
    uint64_t n_prot_global = 0;         // proteins of the whole job
    uint64_t ordinal_base = 0;          // ordinal of this rank's first protein
    uint64_t n_recv = 0;                // records this rank owns after the exchange
    uint64_t exchange_bytes_out = 0;    // bytes of records this rank sent to other ranks in the last build
    uint32_t upload_launches = 0;       // kernels sigk_upload launched (per-protein table, splitters)

    int fail(int code, const char *fmt, ...) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        error = buf;
        return code;
    }
};

#define CU(h, call)                                                                                          \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess)                                                                               \
            return (h)->fail(e_ == cudaErrorMemoryAllocation ? SIGK_E_NOMEM : SIGK_E_CUDA, "%s: %s (%s:%d)", #call, \
                             cudaGetErrorString(e_), __FILE__, __LINE__);                                    \
    } while (0)

namespace sigk {

// (re)size every record-proportional buffer for `cap` records; keep_pingpong1 leaves keys[1]/vals[1] alone
int ensure_capacity(sigk_handle *h, uint64_t cap, bool keep_pingpong1);

// comm.cu
int comm_make_id(void *id128, std::string *err);
int comm_join(sigk_handle *h, const void *id128);
void comm_destroy(sigk_handle *h);
// after upload: learn every rank's protein count, the job-wide max seq_id
int comm_exchange_shapes(sigk_handle *h);
// share the per-protein meta of every rank (h->d_meta holds n_prot_global entries afterwards)
int comm_allgather_meta(sigk_handle *h);
// multi-GPU stage 1: encode + route (+ all-to-all without peer mappings); fills `seg` with the regions, in source-rank
// order, that hold this rank's records (the first sort pass reads them in place)
int comm_encode_exchange(sigk_handle *h, const EncodeArgs &ea, SortSegments *seg, int *first_out, uint32_t *launches);
// min over the ranks of a 0/1 word (with a host synchronisation): collective decisions, and a barrier
int comm_agree(sigk_handle *h, uint64_t mine, uint64_t *out);
// (re)create and map the peer landing zones for the current job size (collective)
int comm_setup_landing(sigk_handle *h);
// once per upload: the splitter codes that cut the k-mer space into one range per rank
int comm_choose_splitters(sigk_handle *h, uint32_t *launches);
// sum over ranks of the per-protein rejected-occurrence counts (before signature_flags)
int comm_reduce_rejected(sigk_handle *h);
// sum the per-rank statistics so that every rank's result carries whole-job counters
int comm_reduce_stats(sigk_handle *h);

}  // namespace sigk
