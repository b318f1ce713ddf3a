// sigk_api.cu — the C ABI of include/sigk.h over the kernels of this directory.
//
// One handle owns one CUDA stream, the device copies of the packed proteins,
// the two record buffers the radix sort ping-pongs between, and the kept-table
// columns.  sigk_build() = extract_kmers + process_kmers of the reference
// (src/kmers-build-signatures.cc:194-196).  No CPU fallback exists: every
// compute entry point needs a CUDA device and fails with SIGK_E_CUDA otherwise.
#include "handle.h"
#include "sigk_common.cuh"

#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace sigk;

namespace {
thread_local std::string g_create_error;

// NVTX range around a stage of the build (header-only NVTX3: a no-op unless a profiler is attached)
struct nvtx_range {
    explicit nvtx_range(const char *name) { nvtxRangePushA(name); }
    ~nvtx_range() { nvtxRangePop(); }
    nvtx_range(const nvtx_range &) = delete;
    nvtx_range &operator=(const nvtx_range &) = delete;
};
}  // namespace

namespace {

int ensure_device(sigk_handle *h) {
    CU(h, cudaSetDevice(h->cfg.device));
    return SIGK_OK;
}

uint16_t *out_col(sigk_handle *h, int c) { return h->d_out_cols.p + (size_t)c * h->capacity; }

}  // namespace

namespace sigk {

int ensure_capacity(sigk_handle *h, uint64_t cap, bool keep_pingpong1) {
    // the multi-GPU path also sorts 16384 samples per rank through the same buffers
    const uint64_t cap_lb = h->comm ? std::max<uint64_t>(cap, 16384ull * 16) : cap;
    for (int i = 0; i < 2; ++i) {
        if (i == 1 && keep_pingpong1) continue;
        CU(h, h->d_keys[i].reserve(cap)); CU(h, h->d_vals[i].reserve(cap));
    }
    // one look-back row set per pass of the main run, one shared by the passes of the side run
    CU(h, h->d_lookback.reserve(onesweep_lookback_bytes(cap_lb) * (SORT_MAX_PASSES + 1)));
    CU(h, h->d_groups.reserve(reduce_group_entries(cap, h->sm_count)));
    CU(h, h->d_long_groups.reserve(reduce_long_group_entries(cap)));
    CU(h, h->d_giant.reserve(reduce_giant_entries(cap)));
    CU(h, h->d_work.reserve(reduce_work_entries(cap, h->sm_count)));
    CU(h, h->d_work_long.reserve(reduce_long_work_entries(cap)));
    CU(h, h->d_rows.reserve(cap));
    CU(h, h->d_out_kmer.reserve(cap));
    CU(h, h->d_out_cols.reserve(cap * 5));
    const uint64_t world = (uint64_t)std::max(1, h->cfg.world);
    CU(h, h->d_scan_state.reserve(std::max<uint64_t>(std::max<uint64_t>((encode_slices(h->total_res) + 1) * world + world, encode_route_state_words(h->total_res)),
                                                     reduce_scan_entries(cap))));
    if (cap > h->capacity) h->capacity = cap;
    return SIGK_OK;
}

}  // namespace sigk

namespace {

int do_upload(sigk_handle *h) {
    if (!h->have_input) return h->fail(SIGK_E_INVALID, "sigk_set_proteins has not been called");
    if (int rc = ensure_device(h)) return rc;
    if (h->cfg.world > 1 && !h->comm) return h->fail(SIGK_E_INVALID, "world > 1: call sigk_comm_join before building");
    const sigk_proteins &p = h->in;
    const uint64_t np = p.n_proteins;
    const uint64_t total = h->total_res;
    const uint64_t padded = encode_tiles(total) * ENC_TILE + ENC_PAD;

    CU(h, h->d_res.reserve(padded));
    CU(h, h->d_starts.reserve(np + 1));
    CU(h, h->d_func.reserve(np));
    CU(h, h->d_seqid.reserve(np));
    CU(h, h->d_slice_prot.reserve(encode_slices(total) + 2));
    CU(h, h->d_hist.reserve(2 * SORT_MAX_PASSES * SIGK_BINS));        // rows 0..7: main run, 8..15: side run
    CU(h, h->d_binbase.reserve(2 * SORT_MAX_PASSES * SIGK_BINS));
    CU(h, h->d_distinct.reserve(SIGK_N_FUNCTION_SLOTS));
    CU(h, h->d_swf.reserve(SIGK_N_FUNCTION_SLOTS));
    CU(h, h->d_scalars.reserve(1));
    CU(h, h->h_scalars.reserve(1));
    CU(h, h->h_distinct.reserve(SIGK_N_FUNCTION_SLOTS));
    CU(h, h->h_swf.reserve(SIGK_N_FUNCTION_SLOTS));
    // one record per residue position is the ceiling (every window valid)
    if (int rc = ensure_capacity(h, std::max<uint64_t>(total, h->capacity), false)) return rc;

    cudaStream_t st = h->stream;
    nvtx_range r_up("sigk upload");
    CU(h, cudaEventRecord(h->ev[EV_START], st));
    if (!h->input_on_device) {          // (sigk_fasta_commit packed the four arrays where they are)
        if (total) CU(h, cudaMemcpyAsync(h->d_res.p, p.residues, total, cudaMemcpyHostToDevice, st));
        CU(h, cudaMemsetAsync(h->d_res.p + total, 0, padded - total, st));
        CU(h, cudaMemcpyAsync(h->d_starts.p, p.starts, (np + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        if (np) {
            CU(h, cudaMemcpyAsync(h->d_func.p, p.function_index, np * sizeof(uint16_t), cudaMemcpyHostToDevice, st));
            CU(h, cudaMemcpyAsync(h->d_seqid.p, p.seq_id, np * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        }
    }
    CU(h, launch_slice_index(h->d_starts.p, (uint32_t)np, total, h->d_slice_prot.p, st));
    CU(h, cudaEventRecord(h->ev[EV_H2D], st));
    CU(h, cudaStreamSynchronize(st));
    cudaEventElapsedTime(&h->h2d_ms, h->ev[EV_START], h->ev[EV_H2D]);
    h->n_prot_global = np;
    h->ordinal_base = 0;
    h->max_seq_id = h->local_max_seq_id;
    h->max_len = h->local_max_len;
    if (h->comm) { if (int rc = comm_exchange_shapes(h)) return rc; }
    // 4-byte table entries when every length fits 16 bits and the 8-byte table would crowd L2 (measured: the
    // unpacking costs 0.9 ms at 2 M proteins where both forms are L2-resident, the smaller table wins 1.5 ms at
    // 8 M proteins).  SIGK_TEST_META=compact|wide overrides the size rule (tests).
    h->meta_compact = h->max_len < 0xFFFFull && h->n_prot_global > (4ull << 20);
    if (const char *force = std::getenv("SIGK_TEST_META")) {
        if (!std::strcmp(force, "compact")) h->meta_compact = h->max_len < 0xFFFFull;
        else if (!std::strcmp(force, "wide")) h->meta_compact = false;
    }
    CU(h, h->d_meta.reserve(meta_bytes(h->n_prot_global, h->meta_compact) << h->meta_shift));
    CU(h, h->d_bitmap.reserve(((uint64_t)h->max_seq_id >> 5) + 1));
    CU(h, h->d_prot_windows.reserve(np));
    CU(h, h->d_prot_rejected.reserve(h->n_prot_global << h->rej_shift));

    // What depends on the input alone is computed here, once per upload, not once per build: the per-protein table
    // (job-wide with a communicator), seqs_with_func (src/signature_build.tcc:160) and the k-mer range splitters.
    uint32_t launches = 0;
    const MetaTable meta{h->d_meta.p, h->meta_compact, h->meta_shift};
    CU(h, cudaMemsetAsync(h->d_swf.p, 0, SIGK_N_FUNCTION_SLOTS * sizeof(uint32_t), st));
    CU(h, launch_protein_meta(h->d_starts.p, h->d_func.p, h->d_seqid.p, (uint32_t)np, meta, h->ordinal_base, h->d_swf.p, st)); ++launches;
    if (h->comm) {
        if (int rc = comm_allgather_meta(h)) return rc;
        if (int rc = comm_choose_splitters(h, &launches)) return rc;
    }
    CU(h, cudaStreamSynchronize(st));
    h->upload_launches = launches;
    h->uploaded = true;
    h->built = h->downloaded = false;
    return SIGK_OK;
}

int do_build_device(sigk_handle *h) {
    if (!h->uploaded) return h->fail(SIGK_E_INVALID, "sigk_upload has not been called");
    if (int rc = ensure_device(h)) return rc;
    cudaStream_t st = h->stream;
    const uint64_t np = h->in.n_proteins;
    DeviceScalars *sc = h->d_scalars.p;
    uint32_t launches = 0;
    nvtx_range r_build("sigk build_device");

    h->table_on_device = false;                 // the table buffers are rewritten from here on
    CU(h, cudaEventRecord(h->ev[EV_DEV0], st));
    CU(h, cudaMemsetAsync(sc, 0, sizeof(DeviceScalars), st));
    CU(h, cudaMemsetAsync(h->d_distinct.p, 0, SIGK_N_FUNCTION_SLOTS * sizeof(uint32_t), st));
    CU(h, cudaMemsetAsync(h->d_bitmap.p, 0, (((uint64_t)h->max_seq_id >> 5) + 1) * sizeof(uint32_t), st));
    CU(h, cudaMemsetAsync(h->d_prot_windows.p, 0, std::max<uint64_t>(np, 1) * sizeof(uint32_t), st));
    CU(h, cudaMemsetAsync(h->d_prot_rejected.p, 0, (std::max<uint64_t>(h->n_prot_global, 1) << h->rej_shift) * sizeof(uint32_t), st));
    CU(h, cudaMemsetAsync(h->d_hist.p, 0, 2 * SORT_MAX_PASSES * SIGK_BINS * sizeof(uint64_t), st));

    // the main run is sorted on the 35 code bits, the side run (records with a lower-case residue) on all 43
    h->plan = make_pass_plan(SIGK_KEY_CODE35_SHIFT, 64);
    h->plan_side = make_pass_plan(SIGK_KEY_CODE_SHIFT, 64);
    const PassPlan &plan = h->plan, &plan_side = h->plan_side;
    if (plan.npass > TK_SIDE0 - TK_SORT0 || plan_side.npass > 16 - TK_SIDE0) return h->fail(SIGK_E_UNSUPPORTED, "too many sort passes");
    // both runs must end in the same ping-pong buffer
    if (((plan.npass - 1) & 1) != (plan_side.npass & 1)) return h->fail(SIGK_E_UNSUPPORTED, "pass counts of the two runs disagree in parity");
    uint64_t *hist_main = h->d_hist.p, *hist_side = h->d_hist.p + (size_t)SORT_MAX_PASSES * SIGK_BINS;
    uint64_t *base_main = h->d_binbase.p, *base_side = h->d_binbase.p + (size_t)SORT_MAX_PASSES * SIGK_BINS;
    const uint64_t cap = h->capacity;
    // (with a communicator the exchange may grow the buffers: the look-back rows are laid out and zeroed after it)
    size_t lb_bytes = onesweep_lookback_bytes(std::max<uint64_t>(cap, 1));
    uint8_t *lb_side = h->d_lookback.p + lb_bytes * SORT_MAX_PASSES;
    if (!h->comm) CU(h, cudaMemsetAsync(h->d_lookback.p, 0, lb_bytes * plan.npass, st));

    const MetaTable meta{h->d_meta.p, h->meta_compact, h->meta_shift};
    EncodeArgs ea{h->d_res.p, h->total_res, h->d_starts.p, (uint32_t)np, (uint32_t)h->ordinal_base, h->d_slice_prot.p, h->d_prot_windows.p};
    int cur;                                    // ping-pong buffer that holds the output of the first pass
    CU(h, cudaEventRecord(h->ev[EV_ENCODE], st));
    CU(h, cudaEventRecord(h->ev[EV_EXCHANGE], st));
    CU(h, cudaEventRecord(h->ev[EV_HIST], st));
    if (!h->comm && h->fused) {
        // ---- stage 1+2a on one GPU: digit counts from the residues, then encode fused with the first radix pass
        nvtx_range r("sigk count + encode-sort");
        CU(h, launch_window_count(ea, plan, hist_main, h->sm_count, st)); ++launches;
        CU(h, launch_scan_bins(hist_main, base_main, plan.npass, &sc->n_main, st)); ++launches;
        CU(h, cudaMemcpyAsync(&sc->n_records, &sc->n_both, sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
        CU(h, cudaEventRecord(h->ev[EV_ENCODE], st));
        CU(h, cudaEventRecord(h->ev[EV_EXCHANGE], st));
        CU(h, cudaEventRecord(h->ev[EV_HIST], st));
        CU(h, cudaEventRecord(h->ev[EV_PASS0], st));
        CU(h, launch_encode_sort(ea, h->d_keys[0].p, h->d_vals[0].p, plan.lo[0], plan.bits[0], base_main, h->d_lookback.p,
                                 sc->ticket + TK_SORT0, h->sm_count, st)); ++launches;
        cur = 0;
    } else {
        // ---- stage 1: encode (+ route to the rank that owns the record's k-mer range); stage 2a: histogram and first
        // pass straight out of the regions the records were written to
        SortSegments seg{};
        const uint64_t *n_ptr = nullptr;
        if (!h->comm) {
            nvtx_range r("sigk encode");
            CU(h, cudaMemsetAsync(h->d_scan_state.p, 0, encode_route_state_words(h->total_res) * sizeof(uint64_t), st));
            EncodeSplitArgs sp{};
            sp.n_split = 0; sp.region_stride = cap; sp.owner_state = h->d_scan_state.p; sp.owner_totals = &sc->n_records;
            sp.overflow = &sc->overflow;
            sp.dst_keys[0] = h->d_keys[0].p; sp.dst_vals[0] = h->d_vals[0].p;
            CU(h, launch_encode_route(ea, sp, sc->ticket + TK_ENCODE, h->sm_count, st)); ++launches;
            seg.n = 1; seg.start[0] = 0; seg.start[1] = cap; seg.keys[0] = h->d_keys[0].p; seg.vals[0] = h->d_vals[0].p;
            n_ptr = &sc->n_records;                 // only the device knows how many windows were valid
            CU(h, cudaEventRecord(h->ev[EV_ENCODE], st));
            CU(h, cudaEventRecord(h->ev[EV_EXCHANGE], st));
            cur = 1;
        } else {
            nvtx_range r("sigk encode + exchange");
            // (records EV_ENCODE; cur = the ping-pong buffer that is free: the regions are the landing zone, or keys[0]
            // after a send/recv exchange)
            if (int rc = comm_encode_exchange(h, ea, &seg, &cur, &launches)) return rc;
            CU(h, cudaEventRecord(h->ev[EV_EXCHANGE], st));
            lb_bytes = onesweep_lookback_bytes(std::max<uint64_t>(h->capacity, 1));
            lb_side = h->d_lookback.p + lb_bytes * SORT_MAX_PASSES;
            CU(h, cudaMemsetAsync(h->d_lookback.p, 0, lb_bytes * plan.npass, st));
        }
        nvtx_range r("sigk histogram + first pass");
        CU(h, launch_histogram_main(seg, n_ptr, plan, hist_main, h->sm_count, st)); ++launches;
        CU(h, launch_scan_bins(hist_main, base_main, plan.npass, &sc->n_main, st)); ++launches;
        CU(h, cudaMemcpyAsync(&sc->n_records, &sc->n_both, sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
        CU(h, cudaEventRecord(h->ev[EV_HIST], st));
        CU(h, cudaEventRecord(h->ev[EV_PASS0], st));
        CU(h, launch_onesweep_first_pass(seg, n_ptr, h->d_keys[cur].p, h->d_vals[cur].p, h->capacity, plan.lo[0], plan.bits[0], base_main,
                                         h->d_lookback.p, sc->ticket + TK_SORT0, h->sm_count, st)); ++launches;
    }
    const uint64_t cap_sort = h->capacity;      // (the exchange may have grown it)

    // ---- stage 2b: the remaining passes of the main run
    {
        nvtx_range r("sigk sort: main run");
        for (int p = 1; p < plan.npass; ++p) {
            CU(h, cudaEventRecord(h->ev[EV_PASS0 + p], st));
            CU(h, launch_onesweep_pass(h->d_keys[cur].p, h->d_vals[cur].p, h->d_keys[cur ^ 1].p, h->d_vals[cur ^ 1].p, &sc->n_main, nullptr,
                                       cap_sort, plan.lo[p], plan.bits[p], base_main + (size_t)p * SIGK_BINS, h->d_lookback.p + lb_bytes * p,
                                       sc->ticket + TK_SORT0 + p, h->sm_count, st)); ++launches;
            cur ^= 1;
        }
        CU(h, cudaEventRecord(h->ev[EV_PASS0 + plan.npass], st));
        CU(h, cudaEventRecord(h->ev[EV_MAIN_SORTED], st));
    }
    // ---- stage 2c: the side run (behind the main run in the first pass's output): its own histogram, all 43 bits
    {
        nvtx_range r("sigk sort: side run");
        int c = cur ^ ((plan.npass - 1) & 1);    // where the first pass left it
        CU(h, launch_histogram(h->d_keys[c].p, &sc->n_side, &sc->n_main, cap_sort, plan_side, hist_side, h->sm_count, st)); ++launches;
        CU(h, launch_scan_bins(hist_side, base_side, plan_side.npass, nullptr, st)); ++launches;
        for (int p = 0; p < plan_side.npass; ++p) {
            CU(h, launch_clear_lookback(lb_side, &sc->n_side, cap_sort, st)); ++launches;
            CU(h, launch_onesweep_pass(h->d_keys[c].p, h->d_vals[c].p, h->d_keys[c ^ 1].p, h->d_vals[c ^ 1].p, &sc->n_side, &sc->n_main,
                                       cap_sort, plan_side.lo[p], plan_side.bits[p], base_side + (size_t)p * SIGK_BINS, lb_side,
                                       sc->ticket + TK_SIDE0 + p, h->sm_count, st)); ++launches;
            c ^= 1;
        }
        if (c != cur) return h->fail(SIGK_E_UNSUPPORTED, "internal: the two sorted runs ended in different buffers");
    }
    CU(h, cudaEventRecord(h->ev[EV_SORT], st));
    h->sorted_in = cur;

    // ---- stages 3+4: run-length + reduce + keep/reject, the order-dependent columns, compaction
    const uint64_t scan_words = reduce_scan_entries(cap_sort);
    CU(h, cudaMemsetAsync(h->d_scan_state.p, 0, scan_words * sizeof(uint64_t), st));
    const int order_stats = (h->cfg.flags & SIGK_F_NO_ORDER_STATS) ? 0 : 1;
    KeptColumns kc{h->d_out_kmer.p, out_col(h, 0), out_col(h, 1), out_col(h, 2), out_col(h, 3), out_col(h, 4)};
    ReduceLists rl{h->d_groups.p, &sc->n_groups, &sc->next_group, h->d_long_groups.p, &sc->n_long, &sc->next_long,
                   h->d_work.p, &sc->n_work, h->d_work_long.p, &sc->n_work_long, h->d_giant.p, &sc->n_giant, &sc->next_giant};
    {
        nvtx_range r("sigk segment reduce");
        CU(h, launch_segment_reduce(h->d_keys[cur].p, h->d_vals[cur].p, &sc->n_records, cap_sort, meta, h->d_rows.p, rl,
                                    h->d_prot_rejected.p, h->d_scan_state.p, &sc->n_segments, order_stats, h->sm_count, st,
                                    h->ev[EV_RED_COUNT], h->ev[EV_RED_EMIT])); launches += 6;
        CU(h, cudaEventRecord(h->ev[EV_REJ0], st));
        if (h->comm) { if (int rc = comm_reduce_rejected(h)) return rc; }
        CU(h, cudaEventRecord(h->ev[EV_REJ1], st));
        CU(h, launch_signature_flags(h->d_prot_windows.p, h->d_prot_rejected.p + h->ordinal_base, h->d_seqid.p, (uint32_t)np, h->d_bitmap.p, st)); ++launches;
        CU(h, launch_popcount(h->d_bitmap.p, ((uint64_t)h->max_seq_id >> 5) + 1, &sc->n_seqs_sig, st)); ++launches;
    }
    CU(h, cudaEventRecord(h->ev[EV_REDUCE], st));
    if (order_stats) {
        nvtx_range r("sigk order statistics");
        CU(h, launch_order_stats(h->d_vals[cur].p, meta, h->d_work.p, &sc->n_work, &sc->next_work, h->d_work_long.p, &sc->n_work_long, &sc->next_work_long, cap_sort, h->d_rows.p, h->sm_count, st)); launches += 2;
    }
    CU(h, cudaEventRecord(h->ev[EV_ORDER], st));
    {
        nvtx_range r("sigk squeeze");
        CU(h, launch_squeeze_rows(h->d_rows.p, &sc->n_segments, cap_sort, kc, h->d_scan_state.p, &sc->n_kept, &sc->n_side_kept, st)); launches += 2;
        // distinct_functions[best]++ per kept row (tcc:286), from the finished function_index column; with a
        // communicator the other ranks' functions are not known here, so both counter ranges are walked
        CU(h, launch_function_histogram(out_col(h, 1), &sc->n_kept, cap_sort, h->comm ? 0xFFFFu : h->local_max_function, h->d_distinct.p,
                                        h->sm_count, st)); ++launches;
        if (h->comm) { if (int rc = comm_reduce_stats(h)) return rc; }
    }
    CU(h, cudaEventRecord(h->ev[EV_SQUEEZE], st));

    h->tm.kernel_launches = launches;
    h->table_on_device = true;
    h->built = true;
    h->downloaded = false;
    return SIGK_OK;
}

int do_download(sigk_handle *h) {
    if (!h->built) return h->fail(SIGK_E_INVALID, "sigk_build_device has not been called");
    if (int rc = ensure_device(h)) return rc;
    cudaStream_t st = h->stream;
    CU(h, cudaMemcpyAsync(h->h_scalars.p, h->d_scalars.p, sizeof(DeviceScalars), cudaMemcpyDeviceToHost, st));
    CU(h, cudaMemcpyAsync(h->h_distinct.p, h->d_distinct.p, SIGK_N_FUNCTION_SLOTS * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU(h, cudaMemcpyAsync(h->h_swf.p, h->d_swf.p, SIGK_N_FUNCTION_SLOTS * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));
    const uint64_t nk = h->h_scalars.p->n_kept;
    if (std::getenv("SIGK_DEBUG_COUNTS")) {
        const DeviceScalars &d = *h->h_scalars.p;
        std::fprintf(stderr, "[sigk] records %llu (main %llu, side %llu) groups %llu kept %llu (side %llu); listed 5..32: %u, longer: %u, walks: %u (+%u whole-warp)\n",
                     (unsigned long long)d.n_records, (unsigned long long)d.n_main, (unsigned long long)d.n_side, (unsigned long long)d.n_segments,
                     (unsigned long long)d.n_kept, (unsigned long long)d.n_side_kept, d.n_groups, d.n_long, d.n_work, d.n_work_long);
    }
    if (nk > h->capacity) return h->fail(SIGK_E_CUDA, "kept count %llu exceeds capacity", (unsigned long long)nk);
    if (nk > h->h_rows) {
        // grow-only pinned result buffers (page-locking is slow; first build pays it)
        CU(h, h->h_kmer.reserve(nk));
        CU(h, h->h_cols.reserve(nk * 5));
        h->h_rows = nk;
    }
    if (nk) {
        CU(h, cudaMemcpyAsync(h->h_kmer.p, h->d_out_kmer.p, nk * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        for (int c = 0; c < 5; ++c)
            CU(h, cudaMemcpyAsync(h->h_cols.p + (size_t)c * h->h_rows, out_col(h, c), nk * sizeof(uint16_t), cudaMemcpyDeviceToHost, st));
    }
    CU(h, cudaEventRecord(h->ev[EV_D2H], st));
    CU(h, cudaStreamSynchronize(st));

    // timings of this build
    sigk_timings &t = h->tm;
    const uint32_t launches = t.kernel_launches;
    std::memset(&t, 0, sizeof t);
    t.kernel_launches = launches;
    auto ms = [&](int a, int b) { float v = 0; cudaEventElapsedTime(&v, h->ev[a], h->ev[b]); return v; };
    t.h2d_ms = h->h2d_ms;
    t.encode_ms = ms(EV_DEV0, EV_ENCODE);
    t.count_ms = (!h->comm && h->fused) ? t.encode_ms : 0.f;
    t.exchange_ms = ms(EV_ENCODE, EV_EXCHANGE);
    t.histogram_ms = ms(EV_EXCHANGE, EV_HIST);
    t.sort_ms = ms(EV_HIST, EV_SORT);
    t.side_sort_ms = ms(EV_MAIN_SORTED, EV_SORT);
    t.reduce_comm_ms = ms(EV_REJ0, EV_REJ1);
    if (h->capacity) {
        t.reduce_count_ms = ms(EV_SORT, EV_RED_COUNT);
        t.reduce_emit_ms = ms(EV_RED_COUNT, EV_RED_EMIT);
        t.reduce_groups_ms = ms(EV_RED_EMIT, EV_REJ0);
    }
    t.reduce_ms = ms(EV_SORT, EV_REDUCE);
    t.order_stats_ms = ms(EV_REDUCE, EV_ORDER);
    t.squeeze_ms = ms(EV_ORDER, EV_SQUEEZE);
    t.d2h_ms = ms(EV_SQUEEZE, EV_D2H);
    t.device_total_ms = ms(EV_DEV0, EV_SQUEEZE);
    t.records_sorted = h->h_scalars.p->n_records;
    t.exchange_bytes_out = h->comm ? h->exchange_bytes_out : 0;
    t.sort_passes = (uint32_t)h->plan.npass;
    t.record_bytes = 12;
    t.key_bytes = 8;
    for (int p = 0; p < h->plan.npass && p < 8; ++p) t.pass_ms[p] = ms(EV_PASS0 + p, EV_PASS0 + p + 1);
    h->downloaded = true;
    return SIGK_OK;
}

}  // namespace

extern "C" {

const char *sigk_version(void) { return "libsigk 0.2 (sm_100a; record 12 B; onesweep, 9-bit digits, 4 passes on 35 code bits, first pass fused with the encode)"; }

int sigk_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int sigk_create(const sigk_config *cfg, sigk_handle **out) {
    if (!cfg || !out) { g_create_error = "null argument"; return SIGK_E_INVALID; }
    *out = nullptr;
    if (cfg->abi_version != SIGK_ABI_VERSION) { g_create_error = "ABI version mismatch"; return SIGK_E_INVALID; }
    if (cfg->k != SIGK_K) { g_create_error = "only K = 8 is supported (src/kmers-build-signatures.cc:17)"; return SIGK_E_UNSUPPORTED; }
    if (cfg->world < 1 || cfg->rank < 0 || cfg->rank >= cfg->world) { g_create_error = "bad rank/world"; return SIGK_E_INVALID; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no usable CUDA device (libsigk has no CPU fallback): ") + cudaGetErrorString(e);
        return SIGK_E_CUDA;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { g_create_error = "device ordinal out of range"; return SIGK_E_INVALID; }
    sigk_handle *h = new sigk_handle;
    h->cfg = *cfg;
    if ((e = cudaSetDevice(cfg->device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        g_create_error = cudaGetErrorString(e);
        delete h;
        return SIGK_E_CUDA;
    }
    for (auto &ev : h->ev) cudaEventCreate(&ev);
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, cfg->device);
    h->fused = std::getenv("SIGK_NO_FUSED") == nullptr;
    if (const char *sp = std::getenv("SIGK_TEST_META_SPREAD")) h->meta_shift = std::max(0, std::min(6, std::atoi(sp)));
    if (const char *sp = std::getenv("SIGK_TEST_REJ_SPREAD")) h->rej_shift = cfg->world > 1 ? 0 : std::max(0, std::min(6, std::atoi(sp)));
    if ((e = onesweep_configure()) != cudaSuccess || (e = reduce_configure(h->meta_shift, h->rej_shift)) != cudaSuccess) {
        g_create_error = std::string("kernel configuration failed (is this an sm_100a device?): ") + cudaGetErrorString(e);
        sigk_destroy(h);
        return SIGK_E_CUDA;
    }
    *out = h;
    return SIGK_OK;
}

void sigk_destroy(sigk_handle *h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    comm_destroy(h);
    h->d_q_res.release(); h->d_q_starts.release(); h->d_q_rows.release();
    h->d_res.release(); h->d_starts.release(); h->d_func.release(); h->d_seqid.release(); h->d_meta.release(); h->d_slice_prot.release();
    for (int i = 0; i < 2; ++i) { h->d_keys[i].release(); h->d_vals[i].release(); }
    h->d_lookback.release(); h->d_hist.release(); h->d_binbase.release(); h->d_scan_state.release();
    h->d_giant.release();
    h->d_groups.release(); h->d_long_groups.release(); h->d_work.release(); h->d_work_long.release(); h->d_rows.release(); h->d_out_kmer.release(); h->d_out_cols.release();
    h->d_bitmap.release(); h->d_distinct.release(); h->d_swf.release(); h->d_scalars.release();
    h->d_prot_windows.release(); h->d_prot_rejected.release();
    h->d_fa_chunk_fn.release(); h->d_fa_chunk_counts.release();
    h->d_fa_bytes.release(); h->d_fa_state.release(); h->d_fa_stream.release(); h->d_fa_tiles.release(); h->d_fa_fn.release(); h->d_fa_err_rec.release();
    h->d_fa_packed.release(); h->d_fa_prefix.release(); h->d_fa_totals.release(); h->d_fa_rec.release(); h->d_fa_err_pos.release(); h->d_fa_src.release();
    h->h_fa_starts.release(); h->h_fa_src.release(); h->h_fa_func.release(); h->h_fa_sid.release();
    h->h_fa_rec.release(); h->h_fa_totals.release(); h->h_fa_err_pos.release(); h->h_fa_err_rec.release();
    h->h_kmer.release(); h->h_cols.release(); h->h_distinct.release(); h->h_swf.release(); h->h_scalars.release();
    for (auto &ev : h->ev) if (ev) cudaEventDestroy(ev);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

const char *sigk_last_error(const sigk_handle *h) { return h ? h->error.c_str() : g_create_error.c_str(); }

void *sigk_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}
void sigk_host_free(void *p) { if (p) cudaFreeHost(p); }

int sigk_set_proteins(sigk_handle *h, const sigk_proteins *p) {
    if (!h) return SIGK_E_INVALID;
    if (!p || !p->starts) return h->fail(SIGK_E_INVALID, "null protein arrays");
    const uint64_t np = p->n_proteins;
    if (np >= 0xFFFFFFFFull) return h->fail(SIGK_E_UNSUPPORTED, "more than 2^32-2 proteins");
    if (p->starts[0] != 0) return h->fail(SIGK_E_INVALID, "starts[0] must be 0");
    const uint64_t total = p->starts[np];
    if (total >= 0xFFFFFFFFull) return h->fail(SIGK_E_UNSUPPORTED, "more than 2^32-2 residues on one GPU");
    if (np && (!p->function_index || !p->seq_id || (total && !p->residues))) return h->fail(SIGK_E_INVALID, "null protein arrays");
    uint32_t max_sid = 0, max_func = 0;
    uint64_t max_len = 0;
    for (uint64_t i = 0; i < np; ++i) {
        if (p->starts[i + 1] < p->starts[i]) return h->fail(SIGK_E_INVALID, "starts must be non-decreasing (protein %llu)", (unsigned long long)i);
        if (p->function_index[i] == SIGK_UNDEFINED_FUNCTION)
            return h->fail(SIGK_E_INVALID, "protein %llu has UndefinedFunction; the host must skip it (src/signature_build.tcc:155)", (unsigned long long)i);
        max_sid = std::max(max_sid, p->seq_id[i]);
        max_func = std::max<uint32_t>(max_func, p->function_index[i]);
        max_len = std::max(max_len, p->starts[i + 1] - p->starts[i]);
    }
    h->local_max_len = max_len;
    h->local_max_function = max_func;
    h->in = *p;
    h->input_on_device = false;
    h->total_res = total;
    h->max_seq_id = max_sid;
    h->local_max_seq_id = max_sid;
    h->have_input = true;
    h->uploaded = h->built = h->downloaded = false;
    return SIGK_OK;
}

int sigk_upload(sigk_handle *h) { return h ? do_upload(h) : SIGK_E_INVALID; }
int sigk_build_device(sigk_handle *h) { return h ? do_build_device(h) : SIGK_E_INVALID; }
int sigk_download(sigk_handle *h) { return h ? do_download(h) : SIGK_E_INVALID; }

int sigk_build(sigk_handle *h) {
    if (!h) return SIGK_E_INVALID;
    if (int rc = do_upload(h)) return rc;
    if (int rc = do_build_device(h)) return rc;
    if (int rc = do_download(h)) return rc;
    return SIGK_OK;
}

int sigk_set_table(sigk_handle *h, const sigk_table *t) {
    if (!h) return SIGK_E_INVALID;
    if (!t || (t->n_kept && !t->kmer)) return h->fail(SIGK_E_INVALID, "null table");
    if (t->n_kept >= 0xFFFFFFFFull) return h->fail(SIGK_E_UNSUPPORTED, "more than 2^32-2 rows");
    if (int rc = ensure_device(h)) return rc;
    cudaStream_t st = h->stream;
    CU(h, h->d_scalars.reserve(1));
    CU(h, h->d_out_kmer.reserve(t->n_kept));
    if (t->n_kept) CU(h, cudaMemcpyAsync(h->d_out_kmer.p, t->kmer, t->n_kept * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    const uint64_t n = t->n_kept;
    CU(h, cudaMemcpyAsync(&h->d_scalars.p->n_kept, &n, sizeof n, cudaMemcpyHostToDevice, st));
    CU(h, cudaStreamSynchronize(st));
    h->table_on_device = true;
    h->built = h->downloaded = false;           // the columns of a previous build no longer match the k-mers
    return SIGK_OK;
}

int sigk_lookup(sigk_handle *h, const uint8_t *residues, const uint64_t *starts, uint64_t n_proteins, uint32_t *rows) {
    if (!h) return SIGK_E_INVALID;
    if (!h->table_on_device) return h->fail(SIGK_E_INVALID, "no table on the device: call sigk_build / sigk_build_device or sigk_set_table first");
    if (!starts || (n_proteins && starts[0] != 0)) return h->fail(SIGK_E_INVALID, "starts[0] must be 0");
    if (n_proteins >= 0xFFFFFFFFull) return h->fail(SIGK_E_UNSUPPORTED, "more than 2^32-2 proteins");
    const uint64_t total = n_proteins ? starts[n_proteins] : 0;
    if (total == 0) return SIGK_OK;
    if (!residues || !rows) return h->fail(SIGK_E_INVALID, "null query arrays");
    if (int rc = ensure_device(h)) return rc;
    cudaStream_t st = h->stream;
    CU(h, h->d_q_res.reserve(total));
    CU(h, h->d_q_starts.reserve(n_proteins + 1));
    CU(h, h->d_q_rows.reserve(total));
    CU(h, cudaMemcpyAsync(h->d_q_res.p, residues, total, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemcpyAsync(h->d_q_starts.p, starts, (n_proteins + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    CU(h, launch_lookup(h->d_q_res.p, h->d_q_starts.p, (uint32_t)n_proteins, total, h->d_out_kmer.p, &h->d_scalars.p->n_kept,
                        h->d_q_rows.p, st));
    CU(h, cudaMemcpyAsync(rows, h->d_q_rows.p, total * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));
    return SIGK_OK;
}

int sigk_result(sigk_handle *h, sigk_table *out) {
    if (!h) return SIGK_E_INVALID;
    if (!out) return h->fail(SIGK_E_INVALID, "null table");
    if (!h->downloaded) return h->fail(SIGK_E_INVALID, "no result: call sigk_build or sigk_download first");
    const DeviceScalars &s = *h->h_scalars.p;
    out->n_kept = s.n_kept;
    out->kmer = reinterpret_cast<const char *>(h->h_kmer.p);
    out->avg_from_end = h->h_cols.p + 0 * h->h_rows;
    out->function_index = h->h_cols.p + 1 * h->h_rows;
    out->mean = h->h_cols.p + 2 * h->h_rows;
    out->median = h->h_cols.p + 3 * h->h_rows;
    out->var = h->h_cols.p + 4 * h->h_rows;
    // with a communicator the counters are whole-job sums; n_kept stays this rank's slice
    out->n_occurrences = h->comm ? s.reduce_in[0] : s.n_records;
    out->n_distinct_kmers = h->comm ? s.reduce_in[1] : s.n_segments;
    out->distinct_signatures = h->comm ? s.reduce_in[2] : s.n_kept;
    out->num_seqs_with_a_signature = s.n_seqs_sig;
    out->distinct_functions = h->h_distinct.p;
    out->seqs_with_func = h->h_swf.p;
    out->n_upper = s.n_kept - s.n_side_kept;
    return SIGK_OK;
}

int sigk_get_timings(const sigk_handle *h, sigk_timings *out) {
    if (!h || !out) return SIGK_E_INVALID;
    *out = h->tm;
    return SIGK_OK;
}

int sigk_synchronize(sigk_handle *h) {
    if (!h) return SIGK_E_INVALID;
    if (int rc = ensure_device(h)) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    return SIGK_OK;
}

int sigk_event_record(sigk_handle *h, int slot) {
    if (!h) return SIGK_E_INVALID;
    if (slot < 0 || slot > 3) return h->fail(SIGK_E_INVALID, "event slot must be 0..3");
    if (int rc = ensure_device(h)) return rc;
    CU(h, cudaEventRecord(h->ev[EV_USER0 + slot], h->stream));
    return SIGK_OK;
}

int sigk_event_elapsed_ms(sigk_handle *h, int slot_a, int slot_b, float *ms) {
    if (!h) return SIGK_E_INVALID;
    if (slot_a < 0 || slot_a > 3 || slot_b < 0 || slot_b > 3 || !ms) return h->fail(SIGK_E_INVALID, "bad event slots");
    CU(h, cudaEventSynchronize(h->ev[EV_USER0 + slot_b]));
    CU(h, cudaEventElapsedTime(ms, h->ev[EV_USER0 + slot_a], h->ev[EV_USER0 + slot_b]));
    return SIGK_OK;
}

int sigk_comm_make_id(void *id128) {
    std::string err;
    const int rc = comm_make_id(id128, &err);
    if (rc) g_create_error = err;
    return rc;
}
int sigk_comm_join(sigk_handle *h, const void *id128) {
    if (!h) return SIGK_E_INVALID;
    h->uploaded = h->built = h->downloaded = false;
    return comm_join(h, id128);
}

int sigk_dbg_encode(sigk_handle *h, const sigk_proteins *p, uint64_t *out_code, uint32_t *out_ordinal,
                    uint16_t *out_offset, uint64_t capacity, uint64_t *n_out) {
    if (!h) return SIGK_E_INVALID;
    if (h->comm) return h->fail(SIGK_E_UNSUPPORTED, "sigk_dbg_encode is single-GPU");
    if (int rc = sigk_set_proteins(h, p)) return rc;
    if (int rc = do_upload(h)) return rc;
    cudaStream_t st = h->stream;
    DeviceScalars *sc = h->d_scalars.p;
    CU(h, cudaMemsetAsync(sc, 0, sizeof(DeviceScalars), st));
    CU(h, cudaMemsetAsync(h->d_scan_state.p, 0, encode_route_state_words(h->total_res) * sizeof(uint64_t), st));
    EncodeArgs ea{h->d_res.p, h->total_res, h->d_starts.p, (uint32_t)p->n_proteins, 0u, h->d_slice_prot.p, nullptr};
    EncodeSplitArgs sp{};
    sp.n_split = 0; sp.region_stride = h->capacity; sp.owner_state = h->d_scan_state.p; sp.owner_totals = &sc->n_records;
    sp.overflow = &sc->overflow;
    sp.dst_keys[0] = h->d_keys[0].p; sp.dst_vals[0] = h->d_vals[0].p;
    CU(h, launch_encode_route(ea, sp, sc->ticket + TK_ENCODE, h->sm_count, st));
    uint64_t n = 0;
    CU(h, cudaMemcpyAsync(&n, &sc->n_records, sizeof n, cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));
    if (n_out) *n_out = n;
    if (n > capacity) return h->fail(SIGK_E_INVALID, "output capacity %llu < %llu records", (unsigned long long)capacity, (unsigned long long)n);
    std::vector<uint64_t> keys(n);
    std::vector<uint32_t> vals(n);
    if (n) {
        CU(h, cudaMemcpy(keys.data(), h->d_keys[0].p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
        CU(h, cudaMemcpy(vals.data(), h->d_vals[0].p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    }
    for (uint64_t i = 0; i < n; ++i) {
        out_code[i] = sigk_key_code(keys[i]);
        out_offset[i] = (uint16_t)sigk_key_offset(keys[i]);
        out_ordinal[i] = vals[i];
    }
    return SIGK_OK;
}

int sigk_dbg_sort_pairs(sigk_handle *h, uint64_t *keys, uint32_t *vals, uint64_t n, int bit_lo, int bit_hi) {
    if (!h) return SIGK_E_INVALID;
    if (bit_lo < 0 || bit_hi > 64 || bit_lo >= bit_hi) return h->fail(SIGK_E_INVALID, "bad bit range");
    if (n >= 0xFFFFFFFFull) return h->fail(SIGK_E_UNSUPPORTED, "too many records");
    if (int rc = ensure_device(h)) return rc;
    if (n == 0) return SIGK_OK;
    cudaStream_t st = h->stream;
    DevBuf<uint64_t> k[2], hist, base, nbuf;
    DevBuf<uint32_t> v[2], ticket;
    DevBuf<uint8_t> lb;
    auto cleanup = [&]() { for (int i = 0; i < 2; ++i) { k[i].release(); v[i].release(); } hist.release(); base.release(); nbuf.release(); ticket.release(); lb.release(); };
    const PassPlan plan = make_pass_plan(bit_lo, bit_hi);
    if (plan.npass > SORT_MAX_PASSES) return h->fail(SIGK_E_INVALID, "too many passes");
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t x) { if (e == cudaSuccess) e = x; return e == cudaSuccess; };
    for (int i = 0; i < 2; ++i) { ok(k[i].reserve(n)); ok(v[i].reserve(n)); }
    ok(hist.reserve(SORT_MAX_PASSES * SIGK_BINS)); ok(base.reserve(SORT_MAX_PASSES * SIGK_BINS));
    ok(nbuf.reserve(1)); ok(ticket.reserve(SORT_MAX_PASSES)); ok(lb.reserve(onesweep_lookback_bytes(n)));
    if (e == cudaSuccess) {
        ok(cudaMemcpyAsync(k[0].p, keys, n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        ok(cudaMemcpyAsync(v[0].p, vals, n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        ok(cudaMemcpyAsync(nbuf.p, &n, sizeof n, cudaMemcpyHostToDevice, st));
        ok(cudaMemsetAsync(hist.p, 0, SORT_MAX_PASSES * SIGK_BINS * sizeof(uint64_t), st));
        ok(cudaMemsetAsync(ticket.p, 0, SORT_MAX_PASSES * sizeof(uint32_t), st));
        ok(launch_histogram(k[0].p, nbuf.p, nullptr, n, plan, hist.p, h->sm_count, st));
        ok(launch_scan_bins(hist.p, base.p, plan.npass, nullptr, st));
        int cur = 0;
        for (int p = 0; p < plan.npass && e == cudaSuccess; ++p) {
            ok(cudaMemsetAsync(lb.p, 0, onesweep_lookback_bytes(n), st));
            ok(launch_onesweep_pass(k[cur].p, v[cur].p, k[cur ^ 1].p, v[cur ^ 1].p, nbuf.p, nullptr, n, plan.lo[p], plan.bits[p],
                                    base.p + (size_t)p * SIGK_BINS, lb.p, ticket.p + p, h->sm_count, st));
            cur ^= 1;
        }
        ok(cudaMemcpyAsync(keys, k[cur].p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        ok(cudaMemcpyAsync(vals, v[cur].p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        ok(cudaStreamSynchronize(st));
    }
    cleanup();
    if (e != cudaSuccess) return h->fail(e == cudaErrorMemoryAllocation ? SIGK_E_NOMEM : SIGK_E_CUDA, "sort: %s", cudaGetErrorString(e));
    return SIGK_OK;
}

int sigk_fasta_parse(sigk_handle *h, const uint8_t *bytes, const uint64_t *file_begin, const uint64_t *file_len, uint64_t n_files,
                     sigk_fasta_records *out) {
    if (!h) return SIGK_E_INVALID;
    if (!out || (n_files && (!bytes || !file_begin || !file_len))) return h->fail(SIGK_E_INVALID, "null argument");
    if (int rc = ensure_device(h)) return rc;
    h->fasta_parsed = false;
    // tiles: FASTA_TILE bytes of one file each
    std::vector<FastaTile> tiles;
    uint64_t span = 0;
    for (uint64_t f = 0; f < n_files; ++f) {
        if (file_begin[f] % 16) return h->fail(SIGK_E_INVALID, "file %llu does not start at a multiple of 16", (unsigned long long)f);
        if (file_begin[f] < span) return h->fail(SIGK_E_INVALID, "files must ascend and not overlap (file %llu)", (unsigned long long)f);
        span = file_begin[f] + file_len[f];
        for (uint64_t off = 0; off < file_len[f]; off += FASTA_TILE)
            tiles.push_back(FastaTile{file_begin[f] + off, (uint32_t)std::min<uint64_t>(FASTA_TILE, file_len[f] - off), off == 0 ? 1u : 0u});
    }
    if (tiles.size() >= 0xFFFFFFFFull) return h->fail(SIGK_E_UNSUPPORTED, "more than 2^32-2 tiles");
    const uint32_t n_tiles = (uint32_t)tiles.size();
    cudaStream_t st = h->stream;
    CU(h, h->d_fa_bytes.reserve(span + 16));
    CU(h, h->d_fa_tiles.reserve(n_tiles));
    CU(h, h->d_fa_fn.reserve(n_tiles));
    CU(h, h->d_fa_chunk_fn.reserve((size_t)n_tiles * FASTA_CHUNKS_PER_TILE));
    CU(h, h->d_fa_chunk_counts.reserve((size_t)n_tiles * FASTA_CHUNKS_PER_TILE));
    CU(h, h->d_fa_state.reserve(n_tiles));
    CU(h, h->d_fa_packed.reserve(n_tiles));
    CU(h, h->d_fa_prefix.reserve(3 * (size_t)n_tiles));
    CU(h, h->d_fa_totals.reserve(3));
    CU(h, h->h_fa_totals.reserve(3));
    nvtx_range r("sigk fasta parse");
    CU(h, cudaEventRecord(h->ev[EV_FA0], st));
    if (span) CU(h, cudaMemcpyAsync(h->d_fa_bytes.p, bytes, span, cudaMemcpyHostToDevice, st));
    if (n_tiles) CU(h, cudaMemcpyAsync(h->d_fa_tiles.p, tiles.data(), n_tiles * sizeof(FastaTile), cudaMemcpyHostToDevice, st));
    CU(h, cudaEventRecord(h->ev[EV_FA1], st));
    CU(h, launch_fasta_tile_functions(h->d_fa_bytes.p, h->d_fa_tiles.p, n_tiles, h->d_fa_chunk_fn.p, h->d_fa_fn.p, h->d_fa_state.p, st));
    CU(h, launch_fasta_count(h->d_fa_bytes.p, h->d_fa_tiles.p, n_tiles, h->d_fa_state.p, h->d_fa_chunk_fn.p, h->d_fa_chunk_counts.p, h->d_fa_packed.p,
                             h->d_fa_prefix.p, h->d_fa_totals.p, st));
    CU(h, cudaMemcpyAsync(h->h_fa_totals.p, h->d_fa_totals.p, 3 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));           // (the tile vector is free to go; the totals size the outputs)
    const uint64_t n_seq = h->h_fa_totals.p[0], n_rec = h->h_fa_totals.p[1], n_err = h->h_fa_totals.p[2];
    const uint64_t err_cap = std::min<uint64_t>(n_err, SIGK_FASTA_MAX_ERRORS);
    const uint64_t stride = n_rec + 1;
    CU(h, h->d_fa_stream.reserve(n_seq + 16));
    CU(h, h->d_fa_rec.reserve(4 * stride));
    CU(h, h->h_fa_rec.reserve(4 * stride));
    CU(h, h->d_fa_err_pos.reserve(err_cap));
    CU(h, h->d_fa_err_rec.reserve(err_cap));
    CU(h, h->h_fa_err_pos.reserve(err_cap));
    CU(h, h->h_fa_err_rec.reserve(err_cap));
    FastaOut o;
    o.residues = h->d_fa_stream.p;
    o.header_pos = h->d_fa_rec.p; o.id_end = h->d_fa_rec.p + stride; o.line_end = h->d_fa_rec.p + 2 * stride; o.seq_begin = h->d_fa_rec.p + 3 * stride;
    o.err_pos = h->d_fa_err_pos.p; o.err_record = h->d_fa_err_rec.p; o.err_capacity = err_cap;
    CU(h, cudaMemsetAsync(o.id_end, 0xFF, 2 * stride * sizeof(uint64_t), st));       // id_end and line_end: "the file ended first"
    CU(h, launch_fasta_emit(h->d_fa_bytes.p, h->d_fa_tiles.p, n_tiles, h->d_fa_state.p, h->d_fa_chunk_fn.p, h->d_fa_chunk_counts.p, h->d_fa_prefix.p, o, st));
    CU(h, cudaMemcpyAsync(o.seq_begin + n_rec, h->d_fa_totals.p, sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
    CU(h, cudaEventRecord(h->ev[EV_FA2], st));
    CU(h, cudaMemcpyAsync(h->h_fa_rec.p, h->d_fa_rec.p, 4 * stride * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    if (err_cap) {
        CU(h, cudaMemcpyAsync(h->h_fa_err_pos.p, h->d_fa_err_pos.p, err_cap * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        CU(h, cudaMemcpyAsync(h->h_fa_err_rec.p, h->d_fa_err_rec.p, err_cap * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    }
    CU(h, cudaEventRecord(h->ev[EV_FA3], st));
    CU(h, cudaStreamSynchronize(st));
    h->fa_records = n_rec; h->fa_residues = n_seq; h->fa_errors = n_err; h->fa_rec_stride = stride; h->fa_bytes = span;
    h->fasta_parsed = true;
    out->n_records = n_rec; out->n_residues = n_seq; out->n_errors = n_err;
    out->header_pos = h->h_fa_rec.p; out->id_end = h->h_fa_rec.p + stride; out->line_end = h->h_fa_rec.p + 2 * stride; out->seq_begin = h->h_fa_rec.p + 3 * stride;
    out->errors = h->h_fa_err_pos.p; out->error_record = h->h_fa_err_rec.p;
    cudaEventElapsedTime(&out->h2d_ms, h->ev[EV_FA0], h->ev[EV_FA1]);
    cudaEventElapsedTime(&out->parse_ms, h->ev[EV_FA1], h->ev[EV_FA2]);
    cudaEventElapsedTime(&out->d2h_ms, h->ev[EV_FA2], h->ev[EV_FA3]);
    return SIGK_OK;
}

int sigk_fasta_commit(sigk_handle *h, const uint8_t *keep, const uint16_t *function_index, const uint32_t *seq_id) {
    if (!h) return SIGK_E_INVALID;
    if (!h->fasta_parsed) return h->fail(SIGK_E_INVALID, "sigk_fasta_parse has not been called");
    const uint64_t n_rec = h->fa_records;
    if (n_rec && (!keep || !function_index || !seq_id)) return h->fail(SIGK_E_INVALID, "null argument");
    if (int rc = ensure_device(h)) return rc;
    const uint64_t *seq_begin = h->h_fa_rec.p + 3 * h->fa_rec_stride;
    uint64_t n_keep = 0;
    for (uint64_t r = 0; r < n_rec; ++r) n_keep += keep[r] != 0;
    if (n_keep >= 0xFFFFFFFFull) return h->fail(SIGK_E_UNSUPPORTED, "more than 2^32-2 proteins");
    // pinned staging for the four small arrays (a pageable source would be copied through the driver's bounce buffer)
    CU(h, h->h_fa_starts.reserve(n_keep + 1));
    CU(h, h->h_fa_src.reserve(n_keep));
    CU(h, h->h_fa_func.reserve(n_keep));
    CU(h, h->h_fa_sid.reserve(n_keep));
    uint64_t *starts = h->h_fa_starts.p, *src = h->h_fa_src.p;
    uint16_t *func = h->h_fa_func.p;
    uint32_t *sid = h->h_fa_sid.p;
    uint32_t max_sid = 0, max_func = 0;
    uint64_t max_len = 0, np = 0;
    starts[0] = 0;
    for (uint64_t r = 0; r < n_rec; ++r) {
        if (!keep[r]) continue;
        if (function_index[r] == SIGK_UNDEFINED_FUNCTION)
            return h->fail(SIGK_E_INVALID, "record %llu has UndefinedFunction; the host must skip it (src/signature_build.tcc:155)", (unsigned long long)r);
        const uint64_t len = seq_begin[r + 1] - seq_begin[r];
        src[np] = seq_begin[r];
        starts[np + 1] = starts[np] + len;
        func[np] = function_index[r];
        sid[np] = seq_id[r];
        max_sid = std::max(max_sid, seq_id[r]);
        max_func = std::max<uint32_t>(max_func, function_index[r]);
        max_len = std::max(max_len, len);
        ++np;
    }
    const uint64_t total = starts[np];
    if (total >= 0xFFFFFFFFull) return h->fail(SIGK_E_UNSUPPORTED, "more than 2^32-2 residues on one GPU");
    const uint64_t padded = encode_tiles(total) * ENC_TILE + ENC_PAD;
    cudaStream_t st = h->stream;
    CU(h, h->d_res.reserve(padded));
    CU(h, h->d_starts.reserve(np + 1));
    CU(h, h->d_func.reserve(np));
    CU(h, h->d_seqid.reserve(np));
    CU(h, h->d_fa_src.reserve(np));
    nvtx_range r("sigk fasta commit");
    CU(h, cudaMemcpyAsync(h->d_starts.p, starts, (np + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    if (np) {
        CU(h, cudaMemcpyAsync(h->d_fa_src.p, src, np * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        CU(h, cudaMemcpyAsync(h->d_func.p, func, np * sizeof(uint16_t), cudaMemcpyHostToDevice, st));
        CU(h, cudaMemcpyAsync(h->d_seqid.p, sid, np * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    }
    CU(h, launch_fasta_gather(h->d_fa_stream.p, h->d_fa_src.p, h->d_starts.p, (uint32_t)np, h->d_res.p, st));
    CU(h, cudaMemsetAsync(h->d_res.p + total, 0, padded - total, st));
    CU(h, cudaStreamSynchronize(st));
    h->in = sigk_proteins{nullptr, nullptr, nullptr, nullptr, np};
    h->input_on_device = true;
    h->local_max_len = max_len;
    h->local_max_function = max_func;
    h->total_res = total;
    h->max_seq_id = max_sid;
    h->local_max_seq_id = max_sid;
    h->have_input = true;
    h->uploaded = h->built = h->downloaded = false;
    return SIGK_OK;
}

int sigk_dbg_fasta_stream(sigk_handle *h, uint8_t *out) {
    if (!h) return SIGK_E_INVALID;
    if (!h->fasta_parsed) return h->fail(SIGK_E_INVALID, "sigk_fasta_parse has not been called");
    if (h->fa_residues) {
        CU(h, cudaMemcpyAsync(out, h->d_fa_stream.p, h->fa_residues, cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
    }
    return SIGK_OK;
}

int sigk_dbg_ddiv(sigk_handle *h, const double *a, const double *b, uint64_t n, double *inl, double *lib) {
    if (!h) return SIGK_E_INVALID;
    if (int rc = ensure_device(h)) return rc;
    if (n == 0) return SIGK_OK;
    cudaStream_t st = h->stream;
    DevBuf<double> d[4];
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t x) { if (e == cudaSuccess) e = x; return e == cudaSuccess; };
    for (auto &x : d) ok(x.reserve(n));
    if (e == cudaSuccess) {
        ok(cudaMemcpyAsync(d[0].p, a, n * sizeof(double), cudaMemcpyHostToDevice, st));
        ok(cudaMemcpyAsync(d[1].p, b, n * sizeof(double), cudaMemcpyHostToDevice, st));
        ok(launch_ddiv_check(d[0].p, d[1].p, n, d[2].p, d[3].p, st));
        ok(cudaMemcpyAsync(inl, d[2].p, n * sizeof(double), cudaMemcpyDeviceToHost, st));
        ok(cudaMemcpyAsync(lib, d[3].p, n * sizeof(double), cudaMemcpyDeviceToHost, st));
        ok(cudaStreamSynchronize(st));
    }
    for (auto &x : d) x.release();
    if (e != cudaSuccess) return h->fail(e == cudaErrorMemoryAllocation ? SIGK_E_NOMEM : SIGK_E_CUDA, "ddiv: %s", cudaGetErrorString(e));
    return SIGK_OK;
}

uint64_t sigk_kmer_encode(const char kmer[8]) {
    uint64_t code = 0, mask = 0;
    for (int i = 0; i < 8; ++i) {
        const int sy = sigk_symbol((unsigned char)kmer[i]);
        if (sy < 0) return UINT64_MAX;
        code = code * 20u + (uint64_t)(sy & 31);
        mask |= (uint64_t)(sy >> 5) << i;
    }
    return (code << SIGK_MASK_BITS) | mask;
}

void sigk_kmer_decode(uint64_t code, char kmer[8]) {
    const uint64_t a = sigk_code_to_ascii(code);
    std::memcpy(kmer, &a, 8);
}

}  // extern "C"
