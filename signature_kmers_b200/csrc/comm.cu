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
    DevBuf<uint64_t> d_samples, d_samples_alt, d_split, d_counts, d_shape;
    DevBuf<uint32_t> d_sample_vals, d_sample_vals_alt, d_bitmaps;
    PinnedBuf<uint64_t> h_counts;
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
    if (c->comm) g_nccl.CommDestroy(c->comm);
    c->d_samples.release(); c->d_samples_alt.release(); c->d_split.release(); c->d_counts.release(); c->d_shape.release();
    c->d_sample_vals.release(); c->d_sample_vals_alt.release(); c->d_bitmaps.release(); c->h_counts.release();
    delete c;
    h->comm = nullptr;
}

int comm_exchange_shapes(sigk_handle *h) {
    Comm *c = h->comm;
    cudaStream_t st = h->stream;
    const int W = c->world;
    CU(h, c->d_shape.reserve(2 + 2 * (size_t)W));
    CU(h, c->h_counts.reserve(std::max<size_t>(2 * (size_t)W, (size_t)W * W)));
    uint64_t mine[2] = {h->in.n_proteins, h->max_seq_id};
    CU(h, cudaMemcpyAsync(c->d_shape.p, mine, sizeof mine, cudaMemcpyHostToDevice, st));
    NC(h, g_nccl.AllGather(c->d_shape.p, c->d_shape.p + 2, 2, ncclUint64, c->comm, st));
    CU(h, cudaMemcpyAsync(c->h_counts.p, c->d_shape.p + 2, 2 * (size_t)W * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));
    c->prot_count.assign(W, 0);
    uint64_t total = 0, base = 0, max_sid = 0;
    for (int r = 0; r < W; ++r) {
        c->prot_count[r] = c->h_counts.p[2 * r];
        if (r < c->rank) base += c->prot_count[r];
        total += c->prot_count[r];
        max_sid = std::max(max_sid, c->h_counts.p[2 * r + 1]);
    }
    if (total >= 0xFFFFFFFFull) return h->fail(SIGK_E_UNSUPPORTED, "more than 2^32-2 proteins in the job");
    h->n_prot_global = total;
    h->ordinal_base = base;
    h->max_seq_id = (uint32_t)max_sid;
    return SIGK_OK;
}

int comm_allgather_meta(sigk_handle *h) {
    Comm *c = h->comm;
    cudaStream_t st = h->stream;
    // every rank's block sits at its ordinal base; one broadcast per rank (blocks differ in size)
    NC(h, g_nccl.GroupStart());
    uint64_t base = 0;
    for (int r = 0; r < c->world; ++r) {
        if (c->prot_count[r])
            NC(h, g_nccl.Broadcast(h->d_meta.p + base, h->d_meta.p + base, c->prot_count[r] * sizeof(uint4), ncclUint8, r, c->comm, st));
        base += c->prot_count[r];
    }
    NC(h, g_nccl.GroupEnd());
    return SIGK_OK;
}

int comm_partition_exchange(sigk_handle *h, uint32_t *launches) {
    Comm *c = h->comm;
    cudaStream_t st = h->stream;
    const int W = c->world;
    DeviceScalars *sc = h->d_scalars.p;
    const uint64_t cap_local = h->total_res;            // records keys[0] can hold so far

    // ---- 1. splitters from sorted samples
    const size_t ns = (size_t)SAMPLES_PER_RANK * W;
    CU(h, c->d_samples.reserve(ns + SAMPLES_PER_RANK)); CU(h, c->d_samples_alt.reserve(ns));
    CU(h, c->d_sample_vals.reserve(ns)); CU(h, c->d_sample_vals_alt.reserve(ns));
    CU(h, c->d_split.reserve(SORT_MAX_SPLIT + 1)); CU(h, c->d_counts.reserve((size_t)W * W + W + 1));
    uint64_t *local = c->d_samples.p + ns;
    sample_keys_kernel<<<(SAMPLES_PER_RANK + 255) / 256, 256, 0, st>>>(h->d_keys[0].p, &sc->n_records, local, SAMPLES_PER_RANK);
    CU(h, cudaGetLastError()); ++*launches;
    NC(h, g_nccl.AllGather(local, c->d_samples.p, SAMPLES_PER_RANK, ncclUint64, c->comm, st));
    {
        // sort the gathered samples on the code bits with the same onesweep kernels
        const PassPlan plan = make_pass_plan(SIGK_KEY_CODE_SHIFT, SIGK_KEY_CODE_SHIFT + SIGK_CODE_BITS);
        uint64_t *nbuf = c->d_counts.p + (size_t)W * W + W;         // holds ns as a device scalar
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
    }

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
    uint64_t n_recv = 0;
    for (int s = 0; s < W; ++s) n_recv += cnt[(size_t)s * W + c->rank];
    if (n_recv >= 0xFFFFFFFFull) return h->fail(SIGK_E_UNSUPPORTED, "more than 2^32-2 records on one rank after the exchange");
    if (int rc = ensure_capacity(h, std::max<uint64_t>(h->capacity, n_recv), /*keep_pingpong1=*/true)) return rc;
    NC(h, g_nccl.GroupStart());
    uint64_t send_off = 0, recv_off = 0;
    for (int r = 0; r < W; ++r) {
        const uint64_t ns_r = cnt[(size_t)c->rank * W + r], nr_r = cnt[(size_t)r * W + c->rank];
        if (ns_r) {
            NC(h, g_nccl.Send(h->d_keys[1].p + send_off, ns_r, ncclUint64, r, c->comm, st));
            NC(h, g_nccl.Send(h->d_vals[1].p + send_off, ns_r, ncclUint32, r, c->comm, st));
        }
        if (nr_r) {
            NC(h, g_nccl.Recv(h->d_keys[0].p + recv_off, nr_r, ncclUint64, r, c->comm, st));
            NC(h, g_nccl.Recv(h->d_vals[0].p + recv_off, nr_r, ncclUint32, r, c->comm, st));
        }
        send_off += ns_r;
        recv_off += nr_r;
    }
    NC(h, g_nccl.GroupEnd());
    // this rank's occurrence count (for the job-wide sum) before n_records becomes the received count
    CU(h, cudaMemcpyAsync(&sc->reduce_in[0], &sc->n_records, sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
    CU(h, cudaMemcpyAsync(&sc->n_records, &n_recv, sizeof n_recv, cudaMemcpyHostToDevice, st));
    CU(h, cudaStreamSynchronize(st));                    // n_recv is a stack variable
    h->n_recv = n_recv;
    // keys[1]/vals[1] are free again: bring them up to the new capacity
    if (int rc = ensure_capacity(h, h->capacity, /*keep_pingpong1=*/false)) return rc;
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
