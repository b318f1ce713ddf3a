// comm.cu — the multi-GPU exchange step: one process per GPU, k-mer space
// range-partitioned over the ranks, one NCCL all-to-all of fixed-size records.
//
// The reference is a single shared-memory process: its
// multimap groups occurrences of a k-mer wherever they were inserted.  Here
// rank r encodes canonical chunk r of the proteins, records are routed to the
// rank that owns their k-mer code range, and every rank then sorts / reduces /
// filters its range on its own.  Ranges are contiguous and ascending in rank, so
// the per-rank kept tables concatenated in rank order are the whole table in
// k-mer order; blocks arrive in source-rank order and the partition pass is
// stable, so equal k-mers keep the canonical insertion order the median / var
// recurrences need.
//
//   1. once per upload: every rank samples SAMPLES_PER_RANK k-mers of its residues; all-gather; sort;
//      splitter k = sample at quantile k/world (same on all ranks); the per-protein table is all-gathered
//      (records carry global ordinals)
//   2. every build: encode_split_kernel stores each record straight into region `source` of its owner's
//      landing zone (peer memory over NVLink, mapped with CUDA IPC); an all-gather of the world x world
//      count matrix tells every rank when all stores have landed and how many records each region holds
//   3. the owner's histogram and first radix pass read the regions in place, in source-rank order
//      (SortSegments): no gather copy
// Without peer mappings the same kernel fills local send regions and grouped ncclSend/ncclRecv moves them.
//
// NCCL is resolved at run time with dlopen("libnccl.so.2") so that a process
// that already loaded a NCCL (torch) shares it and a single-GPU user needs none.
#include "handle.h"
#include "sigk_common.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>

namespace sigk {

namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;

    bool load(std::string *err) {
        if (lib) return true;
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) { if (err) *err = std::string("dlopen libnccl.so.2: ") + dlerror(); return false; }
#define SIGK_SYM(field, name)                                                         \
        *reinterpret_cast<void **>(&field) = dlsym(lib, name);                        \
        if (!field) { if (err) *err = std::string("libnccl.so.2 lacks ") + name; return false; }
        SIGK_SYM(GetUniqueId, "ncclGetUniqueId") SIGK_SYM(CommInitRank, "ncclCommInitRank")
        SIGK_SYM(CommDestroy, "ncclCommDestroy") SIGK_SYM(GetErrorString, "ncclGetErrorString")
        SIGK_SYM(AllGather, "ncclAllGather") SIGK_SYM(AllReduce, "ncclAllReduce") SIGK_SYM(Broadcast, "ncclBroadcast")
        SIGK_SYM(Send, "ncclSend") SIGK_SYM(Recv, "ncclRecv") SIGK_SYM(GroupStart, "ncclGroupStart") SIGK_SYM(GroupEnd, "ncclGroupEnd")
#undef SIGK_SYM
        return true;
    }
};
NcclApi g_nccl;

constexpr int SAMPLES_PER_RANK = 16384;

}  // namespace

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    std::vector<uint64_t> prot_count;          // proteins of every rank
    DevBuf<uint64_t> d_samples, d_samples_alt, d_split, d_counts, d_shape, d_owner_state, d_ipc;
    DevBuf<uint32_t> d_sample_vals, d_sample_vals_alt, d_bitmaps;
    PinnedBuf<uint64_t> h_counts;
    // Peer landing zones: every rank exports one buffer of `world` regions through CUDA IPC; source rank s
    // writes its records for owner d into region s of d's buffer straight from the encode kernel.
    uint64_t *land_keys = nullptr;
    uint32_t *land_vals = nullptr;
    uint64_t land_cap = 0;                      // records the landing zone holds
    uint64_t land_stride = 0;                   // records per (source, owner) region
    uint64_t max_total_res = 0;                 // largest residue count over the ranks
    uint64_t min_stride = 0;                    // region size an overflowing build asked for
    uint64_t *peer_keys[SORT_MAX_SEGMENTS] = {};
    uint32_t *peer_vals[SORT_MAX_SEGMENTS] = {};
    bool peer_ok = false;
};

#define NC(h, call)                                                                                       \
    do {                                                                                                  \
        ncclResult_t r_ = (call);                                                                         \
        if (r_ != ncclSuccess) return (h)->fail(SIGK_E_COMM, "%s: %s", #call, g_nccl.GetErrorString(r_)); \
    } while (0)

namespace {

// Samples for the splitters, taken from the residues: the k-mer at (or after)
// evenly spaced residue positions.  Protein boundaries are ignored — a splitter only has to balance.
__global__ void sample_residues_kernel(const uint8_t *__restrict__ res, uint64_t total_res, uint64_t *__restrict__ samples,
                                       int n_samples) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_samples) return;
    uint64_t key = ~0ull;
    if (total_res >= 8) {
        uint64_t p = (uint64_t)i * (total_res - 7) / n_samples;
        for (int tries = 0; tries < 256 && p + 8 <= total_res; ++tries, ++p) {
            uint64_t code = 0;                   // case-folded: ranges are cut on code35
            bool ok = true;
            for (int j = 0; j < 8; ++j) {
                const int sy = sigk_symbol(res[p + j]);
                if (sy < 0) { ok = false; break; }
                code = code * 20u + (uint64_t)(sy & 31);
            }
            if (ok) { key = sigk_pack_key(code, 0, 0); break; }
        }
    }
    samples[i] = key;
}

__global__ void pick_splitters_kernel(const uint64_t *__restrict__ sorted_samples, int per_rank, int world,
                                      uint64_t *__restrict__ split_codes) {
    const int k = threadIdx.x;      // splitter k = first code owned by rank k+1
    if (k < world - 1) split_codes[k] = sigk_key_code35(sorted_samples[(size_t)(k + 1) * per_rank]);
}

__global__ void or_bitmaps_kernel(const uint32_t *__restrict__ all, uint64_t words, int world, uint32_t *__restrict__ out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t v = 0;
        for (int r = 0; r < world; ++r) v |= all[(uint64_t)r * words + i];
        out[i] = v;
    }
}

}  // namespace

// ---- peer landing zones ---------------------------------------------------------------------------
constexpr size_t IPC_WORDS = 2 * (sizeof(cudaIpcMemHandle_t) / sizeof(uint64_t)) + 1;   // two handles + an ok word

static void close_imports(Comm *c) {
    for (int r = 0; r <= (SORT_MAX_SEGMENTS - 1); ++r) {
        if (r != c->rank) {
            if (c->peer_keys[r]) cudaIpcCloseMemHandle(c->peer_keys[r]);
            if (c->peer_vals[r]) cudaIpcCloseMemHandle(c->peer_vals[r]);
        }
        c->peer_keys[r] = nullptr; c->peer_vals[r] = nullptr;
    }
    c->peer_ok = false;
}

static void release_landing(Comm *c) {
    close_imports(c);
    if (c->land_keys) cudaFree(c->land_keys);
    if (c->land_vals) cudaFree(c->land_vals);
    c->land_keys = nullptr; c->land_vals = nullptr;
    c->land_cap = 0;
}

// min over the ranks of a 0/1 word; doubles as a barrier (every rank's stream has reached this point)
int comm_agree(sigk_handle *h, uint64_t mine, uint64_t *out) {
    Comm *c = h->comm;
    cudaStream_t st = h->stream;
    uint64_t *scratch = c->d_shape.p;
    CU(h, cudaMemcpyAsync(scratch, &mine, sizeof mine, cudaMemcpyHostToDevice, st));
    NC(h, g_nccl.AllReduce(scratch, scratch, 1, ncclUint64, ncclMin, c->comm, st));
    CU(h, cudaMemcpyAsync(c->h_counts.p, scratch, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));
    *out = c->h_counts.p[0];
    return SIGK_OK;
}

// Collective (called from upload with the same max_total_res on every rank): make sure every rank has a
// landing zone of `world` regions big enough for this job and that all of them are mapped here.  Any
// failure on any rank (allocation, IPC export/import, SIGK_NO_PEER_WRITES set) leaves peer_ok false on
// all ranks, and the build routes through local send regions and NCCL send/recv instead.
int comm_setup_landing(sigk_handle *h) {
    Comm *c = h->comm;
    cudaStream_t st = h->stream;
    const int W = c->world;
    const uint64_t per = c->max_total_res / W;
    const uint64_t stride = std::max<uint64_t>((per + per / 4 + 65536 + 63) & ~63ull, c->min_stride);   // expected share + 25 % + slack
    if (c->peer_ok && stride * W <= c->land_cap) { c->land_stride = stride; return SIGK_OK; }
    const bool had = c->land_keys != nullptr;
    if (had) {
        // nobody may still be writing into a zone that is about to go away
        uint64_t dummy;
        if (int rc = comm_agree(h, 1, &dummy)) return rc;
        close_imports(c);
        if (int rc = comm_agree(h, 1, &dummy)) return rc;         // every import is closed before any zone is freed
        release_landing(c);
    }
    uint64_t words[IPC_WORDS] = {};
    bool ok = std::getenv("SIGK_NO_PEER_WRITES") == nullptr;
    if (ok) {
        ok = cudaMalloc(&c->land_keys, stride * W * sizeof(uint64_t)) == cudaSuccess &&
             cudaMalloc(&c->land_vals, stride * W * sizeof(uint32_t)) == cudaSuccess;
        cudaIpcMemHandle_t hk, hv;
        ok = ok && cudaIpcGetMemHandle(&hk, c->land_keys) == cudaSuccess && cudaIpcGetMemHandle(&hv, c->land_vals) == cudaSuccess;
        if (ok) {
            std::memcpy(words, &hk, sizeof hk);
            std::memcpy(words + sizeof hk / sizeof(uint64_t), &hv, sizeof hv);
        }
        cudaGetLastError();
    }
    words[IPC_WORDS - 1] = ok ? 1 : 0;
    CU(h, c->d_ipc.reserve(IPC_WORDS * (size_t)(W + 1)));
    CU(h, cudaMemcpyAsync(c->d_ipc.p + IPC_WORDS * (size_t)W, words, sizeof words, cudaMemcpyHostToDevice, st));
    NC(h, g_nccl.AllGather(c->d_ipc.p + IPC_WORDS * (size_t)W, c->d_ipc.p, IPC_WORDS, ncclUint64, c->comm, st));
    CU(h, cudaMemcpyAsync(c->h_counts.p, c->d_ipc.p, IPC_WORDS * (size_t)W * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));
    bool all = true;
    for (int r = 0; r < W; ++r) all = all && c->h_counts.p[IPC_WORDS * (size_t)r + IPC_WORDS - 1] == 1;
    bool opened = all;
    if (all) {
        for (int r = 0; r < W && opened; ++r) {
            if (r == c->rank) { c->peer_keys[r] = c->land_keys; c->peer_vals[r] = c->land_vals; continue; }
            cudaIpcMemHandle_t hk, hv;
            const uint64_t *w = c->h_counts.p + IPC_WORDS * (size_t)r;
            std::memcpy(&hk, w, sizeof hk);
            std::memcpy(&hv, w + sizeof hk / sizeof(uint64_t), sizeof hv);
            void *pk = nullptr, *pv = nullptr;
            opened = cudaIpcOpenMemHandle(&pk, hk, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
            if (opened) c->peer_keys[r] = static_cast<uint64_t *>(pk);
            opened = opened && cudaIpcOpenMemHandle(&pv, hv, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
            if (opened) c->peer_vals[r] = static_cast<uint32_t *>(pv);
        }
        cudaGetLastError();
    }
    uint64_t everyone = 0;
    if (int rc = comm_agree(h, opened ? 1 : 0, &everyone)) return rc;
    if (everyone == 1) {
        c->peer_ok = true;
        c->land_cap = stride * W;
        c->land_stride = stride;
    } else {
        close_imports(c);
        if (int rc = comm_agree(h, 1, &everyone)) return rc;      // every import is closed before any zone is freed
        release_landing(c);
    }
    return SIGK_OK;
}

int comm_make_id(void *id128, std::string *err) {
    if (!id128) return SIGK_E_INVALID;
    if (!g_nccl.load(err)) return SIGK_E_COMM;
    static_assert(sizeof(ncclUniqueId) == SIGK_COMM_ID_BYTES, "id size");
    ncclUniqueId id;
    const ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) { if (err) *err = g_nccl.GetErrorString(r); return SIGK_E_COMM; }
    std::memcpy(id128, &id, sizeof id);
    return SIGK_OK;
}

int comm_join(sigk_handle *h, const void *id128) {
    if (!id128) return h->fail(SIGK_E_INVALID, "null communicator id");
    if (h->cfg.world < 2) return h->fail(SIGK_E_INVALID, "sigk_comm_join needs world >= 2 in sigk_config");
    if (h->cfg.world > SORT_MAX_SEGMENTS) return h->fail(SIGK_E_UNSUPPORTED, "at most %d ranks", SORT_MAX_SEGMENTS);
    std::string err;
    if (!g_nccl.load(&err)) return h->fail(SIGK_E_COMM, "%s", err.c_str());
    CU(h, cudaSetDevice(h->cfg.device));
    if (h->comm) comm_destroy(h);
    Comm *c = new Comm;
    c->rank = h->cfg.rank;
    c->world = h->cfg.world;
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof id);
    const ncclResult_t r = g_nccl.CommInitRank(&c->comm, c->world, id, c->rank);
    if (r != ncclSuccess) { delete c; return h->fail(SIGK_E_COMM, "ncclCommInitRank: %s", g_nccl.GetErrorString(r)); }
    h->comm = c;
    return SIGK_OK;
}

void comm_destroy(sigk_handle *h) {
    Comm *c = h->comm;
    if (!c) return;
    if (c->peer_ok) {
        // exported memory must not be freed while a peer still has it mapped: close the imports, wait
        // for every rank to have done the same (sigk_destroy is collective when world > 1), then free
        close_imports(c);
        uint64_t dummy = 0;
        comm_agree(h, 1, &dummy);
    }
    release_landing(c);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    c->d_samples.release(); c->d_samples_alt.release(); c->d_split.release(); c->d_counts.release(); c->d_shape.release(); c->d_owner_state.release(); c->d_ipc.release();
    c->d_sample_vals.release(); c->d_sample_vals_alt.release(); c->d_bitmaps.release(); c->h_counts.release();
    delete c;
    h->comm = nullptr;
}

int comm_exchange_shapes(sigk_handle *h) {
    Comm *c = h->comm;
    cudaStream_t st = h->stream;
    const int W = c->world;
    CU(h, c->d_shape.reserve(4 + 4 * (size_t)W));
    CU(h, c->h_counts.reserve(std::max<size_t>(IPC_WORDS * (size_t)W, (size_t)W * (W + 1) + 4)));
    uint64_t mine[4] = {h->in.n_proteins, h->max_seq_id, h->total_res, h->local_max_len};
    CU(h, cudaMemcpyAsync(c->d_shape.p, mine, sizeof mine, cudaMemcpyHostToDevice, st));
    NC(h, g_nccl.AllGather(c->d_shape.p, c->d_shape.p + 4, 4, ncclUint64, c->comm, st));
    CU(h, cudaMemcpyAsync(c->h_counts.p, c->d_shape.p + 4, 4 * (size_t)W * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));
    c->prot_count.assign(W, 0);
    uint64_t total = 0, base = 0, max_sid = 0, max_len = 0;
    c->max_total_res = 0;
    for (int r = 0; r < W; ++r) {
        c->prot_count[r] = c->h_counts.p[4 * r];
        if (r < c->rank) base += c->prot_count[r];
        total += c->prot_count[r];
        max_sid = std::max(max_sid, c->h_counts.p[4 * r + 1]);
        c->max_total_res = std::max(c->max_total_res, c->h_counts.p[4 * r + 2]);
        max_len = std::max(max_len, c->h_counts.p[4 * r + 3]);
    }
    h->max_len = max_len;
    if (total >= 0xFFFFFFFFull) return h->fail(SIGK_E_UNSUPPORTED, "more than 2^32-2 proteins in the job");
    h->n_prot_global = total;
    h->ordinal_base = base;
    h->max_seq_id = (uint32_t)max_sid;
    return comm_setup_landing(h);
}

int comm_allgather_meta(sigk_handle *h) {
    Comm *c = h->comm;
    cudaStream_t st = h->stream;
    // every rank's block sits at its ordinal base; one broadcast per rank (blocks differ in size)
    NC(h, g_nccl.GroupStart());
    uint64_t base = 0;
    for (int r = 0; r < c->world; ++r) {
        if (c->prot_count[r])
            NC(h, g_nccl.Broadcast(h->d_meta.p + (meta_bytes(base, h->meta_compact) << h->meta_shift),
                                   h->d_meta.p + (meta_bytes(base, h->meta_compact) << h->meta_shift),
                                   meta_bytes(c->prot_count[r], h->meta_compact) << h->meta_shift, ncclUint8, r, c->comm, st));
        base += c->prot_count[r];
    }
    NC(h, g_nccl.GroupEnd());
    // seqs_with_func depends on the input alone: summed over the ranks here, once per upload
    NC(h, g_nccl.AllReduce(h->d_swf.p, h->d_swf.p, SIGK_N_FUNCTION_SLOTS, ncclUint32, ncclSum, c->comm, st));
    return SIGK_OK;
}

// All ranks agree on W-1 splitter codes: SAMPLES_PER_RANK k-mers sampled from every rank's residues, all-gathered,
// sorted with the onesweep kernels, cut at the quantiles.  Depends on the input alone: once per upload.
int comm_choose_splitters(sigk_handle *h, uint32_t *launches) {
    Comm *c = h->comm;
    cudaStream_t st = h->stream;
    const int W = c->world;
    DeviceScalars *sc = h->d_scalars.p;
    const size_t ns = (size_t)SAMPLES_PER_RANK * W;
    CU(h, c->d_samples.reserve(ns + SAMPLES_PER_RANK)); CU(h, c->d_samples_alt.reserve(ns));
    CU(h, c->d_sample_vals.reserve(ns)); CU(h, c->d_sample_vals_alt.reserve(ns));
    CU(h, c->d_split.reserve(SORT_MAX_SEGMENTS)); CU(h, c->d_counts.reserve((size_t)W * (W + 1) + W + 8));
    uint64_t *local = c->d_samples.p + ns;
    sample_residues_kernel<<<(SAMPLES_PER_RANK + 255) / 256, 256, 0, st>>>(h->d_res.p, h->total_res, local, SAMPLES_PER_RANK);
    CU(h, cudaGetLastError()); ++*launches;
    NC(h, g_nccl.AllGather(local, c->d_samples.p, SAMPLES_PER_RANK, ncclUint64, c->comm, st));
    // sort the gathered samples on the code bits with the same onesweep kernels
    const PassPlan plan = make_pass_plan(SIGK_KEY_CODE35_SHIFT, 64);
    uint64_t *nbuf = c->d_counts.p + (size_t)W * (W + 1) + W + 4;   // holds ns as a device scalar
    const uint64_t ns64 = ns;
    CU(h, cudaMemcpyAsync(nbuf, &ns64, sizeof ns64, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemsetAsync(h->d_hist.p, 0, SORT_MAX_PASSES * SIGK_BINS * sizeof(uint64_t), st));
    CU(h, launch_histogram(c->d_samples.p, nbuf, nullptr, ns, plan, h->d_hist.p, h->sm_count, st));
    CU(h, launch_scan_bins(h->d_hist.p, h->d_binbase.p, plan.npass, nullptr, st));
    const size_t lb = onesweep_lookback_bytes(ns);
    CU(h, cudaMemsetAsync(h->d_lookback.p, 0, lb * plan.npass, st));
    CU(h, cudaMemsetAsync(sc->ticket, 0, sizeof sc->ticket, st));
    uint64_t *k[2] = {c->d_samples.p, c->d_samples_alt.p};
    uint32_t *v[2] = {c->d_sample_vals.p, c->d_sample_vals_alt.p};
    int cur = 0;
    for (int p = 0; p < plan.npass; ++p) {
        CU(h, launch_onesweep_pass(k[cur], v[cur], k[cur ^ 1], v[cur ^ 1], nbuf, nullptr, ns, plan.lo[p], plan.bits[p],
                                   h->d_binbase.p + (size_t)p * SIGK_BINS, h->d_lookback.p + lb * p,
                                   sc->ticket + TK_SORT0 + p, h->sm_count, st));
        cur ^= 1;
    }
    *launches += 2 + plan.npass;
    pick_splitters_kernel<<<1, 32, 0, st>>>(k[cur], SAMPLES_PER_RANK, W, c->d_split.p);
    CU(h, cudaGetLastError()); ++*launches;
    return SIGK_OK;
}

// what follows every exchange: the local count feeds the job-wide sum (n_records itself comes from the histogram)
static int publish_counts(sigk_handle *h, uint64_t local_records, uint64_t n_recv) {
    Comm *c = h->comm;
    const int W = c->world;
    DeviceScalars *sc = h->d_scalars.p;
    c->h_counts.p[(size_t)W * (W + 1)] = local_records;
    CU(h, cudaMemcpyAsync(&sc->reduce_in[0], c->h_counts.p + (size_t)W * (W + 1), sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
    h->n_recv = n_recv;
    return SIGK_OK;
}

// the all-to-all itself (no peer mappings): region r of keys[1]/vals[1] (at send_base[r]) goes to rank r; what the
// others send lands in keys[0]/vals[0] in source-rank order.  cnt[src * W + dst] is the all-gathered count matrix.
static int exchange_records(sigk_handle *h, const uint64_t *cnt, const uint64_t *send_base) {
    Comm *c = h->comm;
    cudaStream_t st = h->stream;
    const int W = c->world;
    // this rank's own share never leaves the device: a plain copy at HBM speed, outside the NCCL group
    uint64_t recv_base[SORT_MAX_SEGMENTS], recv_off = 0;
    for (int r = 0; r < W; ++r) { recv_base[r] = recv_off; recv_off += cnt[(size_t)r * W + c->rank]; }
    if (const uint64_t own = cnt[(size_t)c->rank * W + c->rank]) {
        CU(h, cudaMemcpyAsync(h->d_keys[0].p + recv_base[c->rank], h->d_keys[1].p + send_base[c->rank], own * sizeof(uint64_t),
                              cudaMemcpyDeviceToDevice, st));
        CU(h, cudaMemcpyAsync(h->d_vals[0].p + recv_base[c->rank], h->d_vals[1].p + send_base[c->rank], own * sizeof(uint32_t),
                              cudaMemcpyDeviceToDevice, st));
    }
    NC(h, g_nccl.GroupStart());
    for (int r = 0; r < W; ++r) {
        if (r == c->rank) continue;
        const uint64_t ns_r = cnt[(size_t)c->rank * W + r], nr_r = cnt[(size_t)r * W + c->rank];
        if (ns_r) {
            NC(h, g_nccl.Send(h->d_keys[1].p + send_base[r], ns_r, ncclUint64, r, c->comm, st));
            NC(h, g_nccl.Send(h->d_vals[1].p + send_base[r], ns_r, ncclUint32, r, c->comm, st));
        }
        if (nr_r) {
            NC(h, g_nccl.Recv(h->d_keys[0].p + recv_base[r], nr_r, ncclUint64, r, c->comm, st));
            NC(h, g_nccl.Recv(h->d_vals[0].p + recv_base[r], nr_r, ncclUint32, r, c->comm, st));
        }
    }
    NC(h, g_nccl.GroupEnd());
    return SIGK_OK;
}

// Multi-GPU stage 1: encode and exchange.  encode_split_kernel writes every record straight into its
// owner's region: with peer mappings that region is in the owner GPU's landing zone (the exchange is the
// kernel's own stores over NVLink, and the owner then sorts straight out of the W regions); without them it
// is a local send region followed by NCCL send/recv.  If a region turns out too small on any rank (skew
// between ranks beyond the 25 % head room), all ranks grow their regions to the exact need and encode again.
// *seg describes where this rank's records are, *first_out which ping-pong buffer the first pass may write.
int comm_encode_exchange(sigk_handle *h, const EncodeArgs &ea, SortSegments *seg, int *first_out, uint32_t *launches) {
    Comm *c = h->comm;
    cudaStream_t st = h->stream;
    const int W = c->world;
    DeviceScalars *sc = h->d_scalars.p;
    uint64_t *totals = c->d_counts.p + (size_t)W * (W + 1);                     // W totals + the overflow flag word
    std::vector<uint64_t> cnt((size_t)W * W);
    for (int attempt = 0;; ++attempt) {
        // every rank's previous build has stopped reading its landing zone before anybody overwrites it: a
        // one-word all-reduce on the build streams is the barrier
        NC(h, g_nccl.AllReduce(c->d_shape.p, c->d_shape.p, 1, ncclUint64, ncclMin, c->comm, st));
        const bool peer = c->peer_ok;
        const uint64_t cap_local = h->total_res;
        // expected share + 25 % + slack, a multiple of 64 records (the kernel stores 16 bytes at a time)
        const uint64_t stride = peer ? c->land_stride
                                     : std::max<uint64_t>((cap_local / W + cap_local / (4 * W) + 65536 + 63) & ~63ull, c->min_stride);
        if (!peer) { CU(h, h->d_keys[1].reserve((size_t)stride * W)); CU(h, h->d_vals[1].reserve((size_t)stride * W)); }
        // SIGK_SPLIT_KERNEL=warp: the round-1 kernel (one run per owner and 512-position warp slice), kept for comparison
        const bool by_tile = !(std::getenv("SIGK_SPLIT_KERNEL") && !std::strcmp(std::getenv("SIGK_SPLIT_KERNEL"), "warp"));
        const size_t state_words = by_tile ? (size_t)encode_route_state_words(h->total_res) : (size_t)encode_slices(h->total_res) * W + W;
        CU(h, c->d_owner_state.reserve(state_words));
        CU(h, cudaMemsetAsync(c->d_owner_state.p, 0, state_words * sizeof(uint64_t), st));
        CU(h, cudaMemsetAsync(totals, 0, (W + 1) * sizeof(uint64_t), st));
        EncodeSplitArgs sp{};
        sp.split_codes = c->d_split.p; sp.n_split = W - 1; sp.region_stride = stride;
        sp.owner_state = c->d_owner_state.p; sp.owner_totals = totals; sp.overflow = reinterpret_cast<uint32_t *>(totals + W);
        for (int d = 0; d < W; ++d) {
            sp.dst_keys[d] = peer ? c->peer_keys[d] + (size_t)c->rank * stride : h->d_keys[1].p + (size_t)d * stride;
            sp.dst_vals[d] = peer ? c->peer_vals[d] + (size_t)c->rank * stride : h->d_vals[1].p + (size_t)d * stride;
        }
        if (by_tile) CU(h, launch_encode_route(ea, sp, sc->ticket + TK_ENCODE, h->sm_count, st));
        else CU(h, launch_encode_split(ea, sp, sc->ticket + TK_ENCODE, st));
        ++*launches;
        CU(h, cudaEventRecord(h->ev[EV_ENCODE], st));
        // W totals + overflow word from every rank; completes only when every rank's kernel (and its peer stores) has
        NC(h, g_nccl.AllGather(totals, c->d_counts.p, W + 1, ncclUint64, c->comm, st));
        CU(h, cudaMemcpyAsync(c->h_counts.p, c->d_counts.p, (size_t)W * (W + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        CU(h, cudaStreamSynchronize(st));
        bool overflow = false;
        uint64_t largest = 0;
        for (int r = 0; r < W; ++r) {
            overflow |= c->h_counts.p[(size_t)r * (W + 1) + W] != 0;
            for (int d = 0; d < W; ++d) {
                cnt[(size_t)r * W + d] = c->h_counts.p[(size_t)r * (W + 1) + d];
                largest = std::max(largest, cnt[(size_t)r * W + d]);
            }
        }
        if (attempt == 0 && std::getenv("SIGK_TEST_FORCE_SPLIT_FALLBACK")) overflow = true;     // tests: exercise the retry on every rank
        // Every rank holds the whole matrix, so every rank reaches the same verdicts below (no rank is left in a collective).
        for (int r = 0; r < W; ++r) {
            uint64_t n_r = 0;
            for (int s2 = 0; s2 < W; ++s2) n_r += cnt[(size_t)s2 * W + r];
            if (n_r >= 0xFFFFFFFFull) return h->fail(SIGK_E_UNSUPPORTED, "more than 2^32-2 records on rank %d after the exchange", r);
        }
        if (overflow) {
            if (attempt >= 2) return h->fail(SIGK_E_COMM, "the exchange regions overflowed three times");
            // the totals are exact even when a region overflowed: size every region for the largest one, with head room
            c->min_stride = (largest + largest / 16 + 4096 + 63) & ~63ull;
            if (int rc = comm_setup_landing(h)) return rc;
            CU(h, cudaMemsetAsync(sc->ticket + TK_ENCODE, 0, sizeof(uint32_t), st));
            if (ea.prot_windows) CU(h, cudaMemsetAsync(ea.prot_windows, 0, std::max<uint64_t>(h->in.n_proteins, 1) * sizeof(uint32_t), st));
            continue;
        }
        uint64_t local = 0, n_recv = 0;
        for (int r = 0; r < W; ++r) { local += cnt[(size_t)c->rank * W + r]; n_recv += cnt[(size_t)r * W + c->rank]; }
        // grow the sort buffers if this rank received more than it encoded; an allocation failure anywhere fails everywhere
        int rc_cap = SIGK_OK;
        if (n_recv > h->capacity) {
            if (!peer) CU(h, cudaStreamSynchronize(st));
            rc_cap = ensure_capacity(h, n_recv, /*keep_pingpong1=*/!peer);
        }
        uint64_t everyone = 0;
        if (int rc = comm_agree(h, rc_cap == SIGK_OK ? 1 : 0, &everyone)) return rc;
        if (!everyone) return rc_cap != SIGK_OK ? rc_cap : h->fail(SIGK_E_NOMEM, "another rank could not grow its sort buffers");
        seg->n = 0;
        if (peer) {
            // everything is already here: the first pass reads the W source regions in place, in source-rank order
            uint64_t off = 0;
            for (int s2 = 0; s2 < W; ++s2) {
                seg->start[s2] = off;
                seg->keys[s2] = c->land_keys + (size_t)s2 * stride;
                seg->vals[s2] = c->land_vals + (size_t)s2 * stride;
                off += cnt[(size_t)s2 * W + c->rank];
            }
            seg->n = W;
            seg->start[W] = off;
            *first_out = 0;
        } else {
            uint64_t send_base[SORT_MAX_SEGMENTS];
            for (int r = 0; r < W; ++r) send_base[r] = (uint64_t)r * stride;
            if (int rc = exchange_records(h, cnt.data(), send_base)) return rc;
            if (h->d_keys[1].cap < h->capacity || h->d_vals[1].cap < h->capacity) {      // the first pass writes into [1]
                CU(h, cudaStreamSynchronize(st));
                if (int rc = ensure_capacity(h, h->capacity, /*keep_pingpong1=*/false)) return rc;
            }
            seg->n = 1;
            seg->start[0] = 0; seg->start[1] = n_recv;
            seg->keys[0] = h->d_keys[0].p; seg->vals[0] = h->d_vals[0].p;
            *first_out = 1;
        }
        h->exchange_bytes_out = 12ull * (local - cnt[(size_t)c->rank * W + c->rank]);
        return publish_counts(h, local, n_recv);
    }
}

int comm_reduce_rejected(sigk_handle *h) {
    Comm *c = h->comm;
    NC(h, g_nccl.AllReduce(h->d_prot_rejected.p, h->d_prot_rejected.p, h->n_prot_global, ncclUint32, ncclSum, c->comm, h->stream));
    return SIGK_OK;
}

int comm_reduce_stats(sigk_handle *h) {
    Comm *c = h->comm;
    cudaStream_t st = h->stream;
    DeviceScalars *sc = h->d_scalars.p;
    // {occurrences, groups, kept} summed over ranks
    CU(h, cudaMemcpyAsync(&sc->reduce_in[1], &sc->n_segments, sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
    CU(h, cudaMemcpyAsync(&sc->reduce_in[2], &sc->n_kept, sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
    NC(h, g_nccl.AllReduce(sc->reduce_in, sc->reduce_in, 3, ncclUint64, ncclSum, c->comm, st));
    NC(h, g_nccl.AllReduce(h->d_distinct.p, h->d_distinct.p, SIGK_N_FUNCTION_SLOTS, ncclUint32, ncclSum, c->comm, st));
    // a protein has a signature if any rank kept one of its k-mers: OR of the bitmaps
    const uint64_t words = ((uint64_t)h->max_seq_id >> 5) + 1;
    CU(h, c->d_bitmaps.reserve(words * c->world));
    NC(h, g_nccl.AllGather(h->d_bitmap.p, c->d_bitmaps.p, words, ncclUint32, c->comm, st));
    or_bitmaps_kernel<<<(unsigned)std::min<uint64_t>((words + 255) / 256, 1024), 256, 0, st>>>(c->d_bitmaps.p, words, c->world, h->d_bitmap.p);
    CU(h, cudaGetLastError());
    CU(h, cudaMemsetAsync(&sc->n_seqs_sig, 0, sizeof(uint64_t), st));
    CU(h, launch_popcount(h->d_bitmap.p, words, &sc->n_seqs_sig, st));
    return SIGK_OK;
}

}  // namespace sigk
