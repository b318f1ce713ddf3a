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
//   1. every rank samples SAMPLES_PER_RANK keys of its encoded records;
//      all-gather; sort; splitter k = sample at quantile k/world (same on all ranks)
//   2. stable multi-way split of the local records by owner (onesweep kernel in
//      SPLIT mode) — after it the records for rank d are one contiguous run
//   3. all-gather the world x world count matrix; grouped ncclSend/ncclRecv
//   4. per-protein meta is all-gathered once per upload (records carry global ordinals)
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
    uint64_t *peer_keys[SORT_MAX_SPLIT + 1] = {};
    uint32_t *peer_vals[SORT_MAX_SPLIT + 1] = {};
    bool peer_ok = false;
};

#define NC(h, call)                                                                                       \
    do {                                                                                                  \
        ncclResult_t r_ = (call);                                                                         \
        if (r_ != ncclSuccess) return (h)->fail(SIGK_E_COMM, "%s: %s", #call, g_nccl.GetErrorString(r_)); \
    } while (0)

namespace {

__global__ void sample_keys_kernel(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ n_ptr,
                                   uint64_t *__restrict__ samples, int n_samples) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_samples) return;
    const uint64_t n = *n_ptr;
    // an empty rank contributes the largest key: it only pulls the top splitter up a little
    samples[i] = n ? keys[(uint64_t)i * n / n_samples] : ~0ull;      // i * n < 2^14 * 2^32
}

// Samples for the fused encode+route path, taken before any record exists: the k-mer at (or after)
// evenly spaced residue positions.  Protein boundaries are ignored — a splitter only has to balance.
__global__ void sample_residues_kernel(const uint8_t *__restrict__ res, uint64_t total_res, uint64_t *__restrict__ samples,
                                       int n_samples) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_samples) return;
    uint64_t key = ~0ull;
    if (total_res >= 8) {
        uint64_t p = (uint64_t)i * (total_res - 7) / n_samples;
        for (int tries = 0; tries < 256 && p + 8 <= total_res; ++tries, ++p) {
            uint64_t code = 0;
            bool ok = true;
            for (int j = 0; j < 8; ++j) {
                const int sy = sigk_symbol(res[p + j]);
                if (sy < 0) { ok = false; break; }
                code = code * 40u + (uint64_t)sy;
            }
            if (ok) { key = sigk_pack_key(code, 0); break; }
        }
    }
    samples[i] = key;
}

__global__ void pick_splitters_kernel(const uint64_t *__restrict__ sorted_samples, int per_rank, int world,
                                      uint64_t *__restrict__ split_codes) {
    const int k = threadIdx.x;      // splitter k = first code owned by rank k+1
    if (k < world - 1) split_codes[k] = sigk_key_code(sorted_samples[(size_t)(k + 1) * per_rank]);
}

__global__ void owner_histogram_kernel(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ n_ptr,
                                       const uint64_t *__restrict__ split_codes, int n_split, uint64_t *__restrict__ counts) {
    __shared__ unsigned long long sh[SORT_MAX_SPLIT + 1];
    __shared__ uint64_t sp[SORT_MAX_SPLIT];
    if (threadIdx.x <= SORT_MAX_SPLIT) sh[threadIdx.x] = 0;
    if ((int)threadIdx.x < n_split) sp[threadIdx.x] = split_codes[threadIdx.x];
    __syncthreads();
    const uint64_t n = *n_ptr;
    uint32_t mine[SORT_MAX_SPLIT + 1] = {};
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t code = sigk_key_code(keys[i]);
        uint32_t d = 0;
        for (int k = 0; k < n_split; ++k) d += code >= sp[k] ? 1u : 0u;
#pragma unroll
        for (int k = 0; k <= SORT_MAX_SPLIT; ++k) mine[k] += (d == (uint32_t)k);
    }
#pragma unroll
    for (int k = 0; k <= SORT_MAX_SPLIT; ++k) {
        uint32_t v = mine[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31u) == 0 && v) atomicAdd(&sh[k], (unsigned long long)v);
    }
    __syncthreads();
    if (threadIdx.x <= SORT_MAX_SPLIT && sh[threadIdx.x]) atomicAdd(reinterpret_cast<unsigned long long *>(counts + threadIdx.x), sh[threadIdx.x]);
}

// bin_base of the partition pass: exclusive scan of the owner counts, padded to SIGK_RADIX entries
__global__ void owner_bases_kernel(const uint64_t *__restrict__ counts, int world, uint64_t *__restrict__ bin_base) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= SIGK_RADIX) return;
    uint64_t run = 0;
    for (int k = 0; k < world && k < d; ++k) run += counts[k];
    bin_base[d] = run;
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
    for (int r = 0; r <= SORT_MAX_SPLIT; ++r) {
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
static int agree(sigk_handle *h, uint64_t mine, uint64_t *out) {
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
static int setup_landing(sigk_handle *h) {
    Comm *c = h->comm;
    cudaStream_t st = h->stream;
    const int W = c->world;
    const uint64_t per = c->max_total_res / W;
    const uint64_t stride = (per + per / 4 + 65536 + 63) & ~63ull;               // expected share + 25 % + slack
    if (c->peer_ok && stride * W <= c->land_cap) { c->land_stride = stride; return SIGK_OK; }
    const bool had = c->land_keys != nullptr;
    if (had) {
        // nobody may still be writing into a zone that is about to go away
        uint64_t dummy;
        if (int rc = agree(h, 1, &dummy)) return rc;
        close_imports(c);
        if (int rc = agree(h, 1, &dummy)) return rc;         // every import is closed before any zone is freed
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
    if (int rc = agree(h, opened ? 1 : 0, &everyone)) return rc;
    if (everyone == 1) {
        c->peer_ok = true;
        c->land_cap = stride * W;
        c->land_stride = stride;
    } else {
        close_imports(c);
        if (int rc = agree(h, 1, &everyone)) return rc;      // every import is closed before any zone is freed
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
    if (h->cfg.world > SORT_MAX_SPLIT + 1) return h->fail(SIGK_E_UNSUPPORTED, "at most %d ranks", SORT_MAX_SPLIT + 1);
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
        agree(h, 1, &dummy);
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
    return setup_landing(h);
}

int comm_allgather_meta(sigk_handle *h) {
    Comm *c = h->comm;
    cudaStream_t st = h->stream;
    // every rank's block sits at its ordinal base; one broadcast per rank (blocks differ in size)
    NC(h, g_nccl.GroupStart());
    uint64_t base = 0;
    for (int r = 0; r < c->world; ++r) {
        if (c->prot_count[r])
            NC(h, g_nccl.Broadcast(h->d_meta.p + meta_bytes(base, h->meta_compact), h->d_meta.p + meta_bytes(base, h->meta_compact),
                                   meta_bytes(c->prot_count[r], h->meta_compact), ncclUint8, r, c->comm, st));
        base += c->prot_count[r];
    }
    NC(h, g_nccl.GroupEnd());
    return SIGK_OK;
}

// All ranks agree on W-1 splitter codes: SAMPLES_PER_RANK local samples (of the encoded keys, or of the
// residues when nothing is encoded yet), all-gathered, sorted with the onesweep kernels, cut at the quantiles.
static int choose_splitters(sigk_handle *h, bool from_residues, uint32_t *launches) {
    Comm *c = h->comm;
    cudaStream_t st = h->stream;
    const int W = c->world;
    DeviceScalars *sc = h->d_scalars.p;
    const size_t ns = (size_t)SAMPLES_PER_RANK * W;
    CU(h, c->d_samples.reserve(ns + SAMPLES_PER_RANK)); CU(h, c->d_samples_alt.reserve(ns));
    CU(h, c->d_sample_vals.reserve(ns)); CU(h, c->d_sample_vals_alt.reserve(ns));
    CU(h, c->d_split.reserve(SORT_MAX_SPLIT + 1)); CU(h, c->d_counts.reserve((size_t)W * (W + 1) + W + 8));
    uint64_t *local = c->d_samples.p + ns;
    if (from_residues)
        sample_residues_kernel<<<(SAMPLES_PER_RANK + 255) / 256, 256, 0, st>>>(h->d_res.p, h->total_res, local, SAMPLES_PER_RANK);
    else
        sample_keys_kernel<<<(SAMPLES_PER_RANK + 255) / 256, 256, 0, st>>>(h->d_keys[0].p, &sc->n_records, local, SAMPLES_PER_RANK);
    CU(h, cudaGetLastError()); ++*launches;
    NC(h, g_nccl.AllGather(local, c->d_samples.p, SAMPLES_PER_RANK, ncclUint64, c->comm, st));
    // sort the gathered samples on the code bits with the same onesweep kernels
    const PassPlan plan = make_pass_plan(SIGK_KEY_CODE_SHIFT, SIGK_KEY_CODE_SHIFT + SIGK_CODE_BITS);
    uint64_t *nbuf = c->d_counts.p + (size_t)W * (W + 1) + W + 4;   // holds ns as a device scalar
    const uint64_t ns64 = ns;
    CU(h, cudaMemcpyAsync(nbuf, &ns64, sizeof ns64, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemsetAsync(h->d_hist.p, 0, SORT_MAX_PASSES * SIGK_RADIX * sizeof(uint64_t), st));
    CU(h, launch_histogram(c->d_samples.p, nbuf, ns, plan, h->d_hist.p, h->sm_count, st));
    CU(h, launch_scan_bins(h->d_hist.p, h->d_binbase.p, plan.npass, st));
    const size_t lb = onesweep_lookback_bytes(ns);
    CU(h, cudaMemsetAsync(h->d_lookback.p, 0, lb * plan.npass, st));
    CU(h, cudaMemsetAsync(sc->ticket + TK_SORT0, 0, SORT_MAX_PASSES * sizeof(uint32_t), st));
    uint64_t *k[2] = {c->d_samples.p, c->d_samples_alt.p};
    uint32_t *v[2] = {c->d_sample_vals.p, c->d_sample_vals_alt.p};
    int cur = 0;
    for (int p = 0; p < plan.npass; ++p) {
        CU(h, launch_onesweep_pass(k[cur], v[cur], k[cur ^ 1], v[cur ^ 1], nbuf, ns, plan.lo[p], plan.bits[p],
                                   h->d_binbase.p + (size_t)p * SIGK_RADIX, h->d_lookback.p + lb * p,
                                   sc->ticket + TK_SORT0 + p, st));
        cur ^= 1;
    }
    *launches += 2 + plan.npass;
    pick_splitters_kernel<<<1, 32, 0, st>>>(k[cur], SAMPLES_PER_RANK, W, c->d_split.p);
    CU(h, cudaGetLastError()); ++*launches;
    CU(h, cudaMemsetAsync(sc->ticket + TK_SORT0, 0, SORT_MAX_PASSES * sizeof(uint32_t), st));
    CU(h, cudaMemsetAsync(h->d_hist.p, 0, SORT_MAX_PASSES * SIGK_RADIX * sizeof(uint64_t), st));
    return SIGK_OK;
}

static int publish_counts(sigk_handle *h, uint64_t local_records, uint64_t n_recv);

// the all-to-all itself: region r of keys[1]/vals[1] (at send_base[r]) goes to rank r; what the others send
// lands in keys[0]/vals[0] in source-rank order.  cnt[src * W + dst] is the all-gathered count matrix.
static int exchange_records(sigk_handle *h, const uint64_t *cnt, const uint64_t *send_base, uint64_t local_records) {
    Comm *c = h->comm;
    cudaStream_t st = h->stream;
    const int W = c->world;
    DeviceScalars *sc = h->d_scalars.p;
    uint64_t n_recv = 0;
    for (int s2 = 0; s2 < W; ++s2) n_recv += cnt[(size_t)s2 * W + c->rank];
    if (n_recv >= 0xFFFFFFFFull) return h->fail(SIGK_E_UNSUPPORTED, "more than 2^32-2 records on one rank after the exchange");
    if (int rc = ensure_capacity(h, std::max<uint64_t>(h->capacity, n_recv), /*keep_pingpong1=*/true)) return rc;
    // this rank's own share never leaves the device: a plain copy at HBM speed, outside the NCCL group
    uint64_t recv_base[SORT_MAX_SPLIT + 1], recv_off = 0;
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
    return publish_counts(h, local_records, n_recv);
}

int comm_partition_exchange(sigk_handle *h, uint32_t *launches) {
    Comm *c = h->comm;
    cudaStream_t st = h->stream;
    const int W = c->world;
    DeviceScalars *sc = h->d_scalars.p;
    const uint64_t cap_local = h->total_res;            // records keys[0] can hold so far

    // ---- 1. splitters from sorted samples of the encoded keys
    if (int rc = choose_splitters(h, /*from_residues=*/false, launches)) return rc;

    // ---- 2. stable split of the local records by owner: keys[0] -> keys[1]
    uint64_t *my_counts = c->d_counts.p + (size_t)W * W;
    CU(h, cudaMemsetAsync(my_counts, 0, W * sizeof(uint64_t), st));
    owner_histogram_kernel<<<h->sm_count * 4, 256, 0, st>>>(h->d_keys[0].p, &sc->n_records, c->d_split.p, W - 1, my_counts);
    CU(h, cudaGetLastError());
    owner_bases_kernel<<<(SIGK_RADIX + 255) / 256, 256, 0, st>>>(my_counts, W, h->d_binbase.p);
    CU(h, cudaGetLastError());
    CU(h, cudaMemsetAsync(h->d_lookback.p, 0, onesweep_lookback_bytes(cap_local), st));
    CU(h, launch_onesweep_partition(h->d_keys[0].p, h->d_vals[0].p, h->d_keys[1].p, h->d_vals[1].p, &sc->n_records, cap_local,
                                    c->d_split.p, W - 1, h->d_binbase.p, h->d_lookback.p, sc->ticket + TK_PARTITION, st));
    *launches += 3;

    // ---- 3. counts matrix, then the all-to-all
    NC(h, g_nccl.AllGather(my_counts, c->d_counts.p, W, ncclUint64, c->comm, st));
    CU(h, cudaMemcpyAsync(c->h_counts.p, c->d_counts.p, (size_t)W * W * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));
    const uint64_t *cnt = c->h_counts.p;                 // cnt[src * W + dst]
    uint64_t send_base[SORT_MAX_SPLIT + 1], run = 0;
    for (int r = 0; r < W; ++r) { send_base[r] = run; run += cnt[(size_t)c->rank * W + r]; }
    if (int rc = exchange_records(h, cnt, send_base, run)) return rc;
    CU(h, cudaStreamSynchronize(st));                    // keys[1]/vals[1] are free again: bring them up to the new capacity
    if (int rc = ensure_capacity(h, h->capacity, /*keep_pingpong1=*/false)) return rc;
    return SIGK_OK;
}

// what follows every exchange: the received count becomes n_records, the local count feeds the job-wide sum
static int publish_counts(sigk_handle *h, uint64_t local_records, uint64_t n_recv) {
    Comm *c = h->comm;
    const int W = c->world;
    DeviceScalars *sc = h->d_scalars.p;
    c->h_counts.p[(size_t)W * (W + 1)] = local_records;
    c->h_counts.p[(size_t)W * (W + 1) + 1] = n_recv;
    CU(h, cudaMemcpyAsync(&sc->reduce_in[0], c->h_counts.p + (size_t)W * (W + 1), sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(&sc->n_records, c->h_counts.p + (size_t)W * (W + 1) + 1, sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
    h->n_recv = n_recv;
    return SIGK_OK;
}

// Multi-GPU stage 1: encode and exchange.  encode_split_kernel writes every record straight into its
// owner's region: with peer mappings that region is in the owner GPU's landing zone (the exchange is the
// kernel's own stores over NVLink, and what remains is a local gather of the W regions into sort order);
// without them it is a local send region followed by NCCL send/recv.  If a region turns out too small on any
// rank (pathological skew between ranks), all ranks fall back to encode + stable split pass + send/recv.
int comm_encode_exchange(sigk_handle *h, const EncodeArgs &ea, uint32_t *launches) {
    Comm *c = h->comm;
    cudaStream_t st = h->stream;
    const int W = c->world;
    DeviceScalars *sc = h->d_scalars.p;
    // the all-gather in here is also the barrier that makes the landing zones safe to overwrite: it completes
    // only after every rank's stream has finished the previous build
    if (int rc = choose_splitters(h, /*from_residues=*/true, launches)) return rc;

    const bool peer = c->peer_ok;
    const uint64_t cap_local = h->total_res;
    // expected share + 25 % + slack, a multiple of 64 records (the kernel stores 16 bytes at a time)
    const uint64_t stride = peer ? c->land_stride : (cap_local / W + cap_local / (4 * W) + 65536 + 63) & ~63ull;
    if (!peer) { CU(h, h->d_keys[1].reserve((size_t)stride * W)); CU(h, h->d_vals[1].reserve((size_t)stride * W)); }
    const size_t state_words = (size_t)encode_slices(h->total_res) * W + W;
    CU(h, c->d_owner_state.reserve(state_words));
    CU(h, cudaMemsetAsync(c->d_owner_state.p, 0, state_words * sizeof(uint64_t), st));
    uint64_t *totals = c->d_counts.p + (size_t)W * (W + 1);                     // W totals + the overflow flag word
    CU(h, cudaMemsetAsync(totals, 0, (W + 1) * sizeof(uint64_t), st));
    EncodeSplitArgs sp{};
    sp.split_codes = c->d_split.p; sp.n_split = W - 1; sp.region_stride = stride;
    sp.owner_state = c->d_owner_state.p; sp.owner_totals = totals; sp.overflow = reinterpret_cast<uint32_t *>(totals + W);
    for (int d = 0; d < W; ++d) {
        sp.dst_keys[d] = peer ? c->peer_keys[d] + (size_t)c->rank * stride : h->d_keys[1].p + (size_t)d * stride;
        sp.dst_vals[d] = peer ? c->peer_vals[d] + (size_t)c->rank * stride : h->d_vals[1].p + (size_t)d * stride;
    }
    CU(h, launch_encode_split(ea, sp, sc->ticket + TK_ENCODE, st)); ++*launches;
    CU(h, cudaEventRecord(h->ev[EV_ENCODE], st));
    // W totals + overflow word from every rank; completes only when every rank's kernel (and its peer stores) has
    NC(h, g_nccl.AllGather(totals, c->d_counts.p, W + 1, ncclUint64, c->comm, st));
    CU(h, cudaMemcpyAsync(c->h_counts.p, c->d_counts.p, (size_t)W * (W + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));
    bool overflow = false;
    std::vector<uint64_t> cnt((size_t)W * W);
    for (int r = 0; r < W; ++r) {
        overflow |= c->h_counts.p[(size_t)r * (W + 1) + W] != 0;
        for (int d = 0; d < W; ++d) cnt[(size_t)r * W + d] = c->h_counts.p[(size_t)r * (W + 1) + d];
    }
    if (std::getenv("SIGK_TEST_FORCE_SPLIT_FALLBACK")) overflow = true;     // tests: exercise the fallback on every rank
    if (overflow) {
        CU(h, cudaMemsetAsync(sc->ticket + TK_ENCODE, 0, sizeof(uint32_t), st));
        CU(h, cudaMemsetAsync(h->d_prot_windows.p, 0, std::max<uint64_t>(h->in.n_proteins, 1) * sizeof(uint32_t), st));
        CU(h, cudaMemsetAsync(h->d_scan_state.p, 0, encode_scan_entries(h->total_res) * sizeof(uint64_t), st));
        CU(h, launch_encode(ea, h->d_keys[0].p, h->d_vals[0].p, h->d_scan_state.p, sc->ticket + TK_ENCODE, &sc->n_records, st)); ++*launches;
        return comm_partition_exchange(h, launches);
    }
    uint64_t local = 0;
    for (int r = 0; r < W; ++r) local += cnt[(size_t)c->rank * W + r];
    if (peer) {
        // everything is already here: gather the W source regions, in source-rank order, into the sort input
        uint64_t n_recv = 0;
        for (int s2 = 0; s2 < W; ++s2) n_recv += cnt[(size_t)s2 * W + c->rank];
        if (n_recv >= 0xFFFFFFFFull) return h->fail(SIGK_E_UNSUPPORTED, "more than 2^32-2 records on one rank after the exchange");
        if (int rc = ensure_capacity(h, std::max<uint64_t>(h->capacity, n_recv), /*keep_pingpong1=*/false)) return rc;
        uint64_t off = 0;
        for (int s2 = 0; s2 < W; ++s2) {
            const uint64_t n = cnt[(size_t)s2 * W + c->rank];
            if (n) {
                CU(h, cudaMemcpyAsync(h->d_keys[0].p + off, c->land_keys + (size_t)s2 * stride, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
                CU(h, cudaMemcpyAsync(h->d_vals[0].p + off, c->land_vals + (size_t)s2 * stride, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
            }
            off += n;
        }
        return publish_counts(h, local, n_recv);
    }
    uint64_t send_base[SORT_MAX_SPLIT + 1];
    for (int r = 0; r < W; ++r) send_base[r] = (uint64_t)r * stride;
    if (int rc = exchange_records(h, cnt.data(), send_base, local)) return rc;
    if (h->d_keys[1].cap < h->capacity || h->d_vals[1].cap < h->capacity) {      // the sort ping-pongs through [1]
        CU(h, cudaStreamSynchronize(st));
        if (int rc = ensure_capacity(h, h->capacity, /*keep_pingpong1=*/false)) return rc;
    }
    return SIGK_OK;
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
    NC(h, g_nccl.AllReduce(h->d_swf.p, h->d_swf.p, SIGK_N_FUNCTION_SLOTS, ncclUint32, ncclSum, c->comm, st));
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
