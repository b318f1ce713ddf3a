// lookup.cu — batch k-mer lookups against the kept table that is resident on the device
// (the consumer side of the table).
//
// Replaces, for a whole batch of query proteins at once, the per-window
//   for_each_kmer(...)             src/kmer_data.h:76-102
//   KeptKmerDB<K>::fetch(kmer,...) src/kept_kmer_db.h:20-28   (exact membership)
// pair that FunctionCaller::process_aa_seq drives (src/call_functions.tcc:262-343).  One thread per
// residue position: the window starting there is visited iff it lies inside its protein and no '*' / 'X'
// sits in the window or right behind it (`kend >= next_ambig`, src/kmer_data.h:90 — the sequential scan
// visits exactly those positions); the 8 raw bytes (case preserved, no alphabet test on the call side) are
// looked up in the table's k-mer column by binary search under the table's order (include/sigk.h): k-mers
// without a lower-case residue first, in byte order, then the others by (case-folded bytes, case mask).  The
// order is defined on every 8-byte string (fold = clear bit 0x20 of each byte, mask = those bits), so a query
// with bytes no table row can hold simply is not found.
// rows[g] = table row of the window at residue position g, or 0xFFFFFFFF.
//
// HBM traffic per position: 1 residue byte read, 4 bytes written, and a binary search whose upper levels
// stay in L2 (log2(rows) dependent 8-byte reads; about the last 8 levels miss for a 371 M-row table).
#include "kernels.h"
#include "sigk_common.cuh"

namespace sigk {

namespace {

SIGK_D uint64_t bswap64(uint64_t x) {
    const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    return ((uint64_t)__byte_perm(lo, 0, 0x0123) << 32) | (uint64_t)__byte_perm(hi, 0, 0x0123);
}

// a < b in table order; a and b are k-mers as big-endian integers (first residue in the top byte)
SIGK_D bool table_less(uint64_t a, uint64_t b) {
    constexpr uint64_t CASE = 0x2020202020202020ull;
    const uint64_t ma = a & CASE, mb = b & CASE;
    if ((ma != 0) != (mb != 0)) return mb != 0;               // section: no lower-case residue first
    const uint64_t fa = a & ~CASE, fb = b & ~CASE;
    if (fa != fb) return fa < fb;
    // case mask with residue j in bit j: the first residue is the LEAST significant bit, i.e. the byte order of
    // the big-endian mask word reversed
    return bswap64(ma) < bswap64(mb);
}

__global__ void __launch_bounds__(256)
lookup_kernel(const uint8_t *__restrict__ res, const uint64_t *__restrict__ starts, uint32_t n_prot, uint64_t total,
              const uint64_t *__restrict__ table_kmers, const uint64_t *__restrict__ n_rows_ptr, uint32_t *__restrict__ rows) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    // protein holding position g: last i with starts[i] <= g
    uint32_t lo = 0, hi = n_prot;                       // starts[lo] <= g < starts[hi]
    while (hi - lo > 1) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (__ldg(starts + mid) <= g) lo = mid; else hi = mid;
    }
    const uint64_t end = __ldg(starts + lo + 1);
    uint32_t row = 0xFFFFFFFFu;
    if (g + SIGK_K_DEV <= end) {
        uint64_t key = 0;
        bool ok = true;
#pragma unroll
        for (int j = 0; j < SIGK_K_DEV; ++j) {
            const uint8_t c = res[g + j];
            ok = ok && c != '*' && c != 'X';
            key = (key << 8) | c;
        }
        if (g + SIGK_K_DEV < end) {                     // the character right behind the window
            const uint8_t c = res[g + SIGK_K_DEV];
            ok = ok && c != '*' && c != 'X';
        }
        if (ok) {
            uint64_t a = 0, b = *n_rows_ptr;            // first row >= key
            while (a < b) {
                const uint64_t mid = a + (b - a) / 2;
                if (table_less(bswap64(__ldg(table_kmers + mid)), key)) a = mid + 1; else b = mid;
            }
            if (a < *n_rows_ptr && bswap64(__ldg(table_kmers + a)) == key) row = (uint32_t)a;
        }
    }
    rows[g] = row;
}

}  // namespace

cudaError_t launch_lookup(const uint8_t *res, const uint64_t *starts, uint32_t n_prot, uint64_t total, const uint64_t *table_kmers,
                          const uint64_t *n_rows_ptr, uint32_t *rows, cudaStream_t stream) {
    if (total == 0 || n_prot == 0) return cudaSuccess;
    lookup_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(res, starts, n_prot, total, table_kmers, n_rows_ptr, rows);
    return cudaGetLastError();
}

}  // namespace sigk
