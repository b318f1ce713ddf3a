"""Pure-Python restatement of the signature-generation hot path, written
independently of oracle/sigk_oracle.cpp so the two can check each other.

TEST INFRASTRUCTURE ONLY (see oracle/sigk_oracle.h); small inputs only.
PARITY: pinned through oracle/sigk_oracle.cpp (see sigk_oracle.h); UNPINNED against a stock reference run for the
Boost/TBB internals — the reference has no golden vectors and cannot be built as a whole here.

Follows the reference's src/signature_build.tcc:
  load_kmers_from_sequence  :120-181
  process_kmers             :183-213
  process_kmer_set          :218-293
with Boost.Accumulators (sum/mean/median=P^2/variance) and the TBB multimap's
newest-first order inside a key restated from their published algorithms.
Python floats are IEEE-754 doubles with every operation rounded once, which is
what g++ -O3 without -march produces for the reference.
"""
from __future__ import annotations

import bisect
import math

import numpy as np

K = 8
OK_PROT = set(b"ACDEFGHIKLMNPQRSTVWYacdefghiklmnpqrstvwy")  # src/signature_build.h:102-103
UNDEFINED = 0xFFFF


def u16_from_double(d: float) -> int:
    """(unsigned short)double via cvttsd2si on x86-64."""
    if d != d or d >= 2147483648.0 or d <= -2147483649.0:
        i = -(1 << 31)
    else:
        i = int(d)  # truncates toward zero
    return i & 0xFFFF


class BoostAcc:
    """accumulator_set<unsigned short, stats<mean, median, variance>> (:262-264)."""

    INC = (0.0, 0.25, 0.5, 0.75, 1.0)

    def __init__(self):
        self.n = 0
        self.S = 0  # unsigned short running sum
        self.var = 0.0
        self.q = [0.0] * 5
        self.pos = [1.0, 2.0, 3.0, 4.0, 5.0]
        self.des = [1.0, 2.0, 3.0, 4.0, 5.0]

    def push(self, x: int):
        self.n += 1
        n = self.n
        self.S = (self.S + x) & 0xFFFF
        q, pos, des = self.q, self.pos, self.des
        if n <= 5:
            q[n - 1] = float(x)
            if n == 5:
                q.sort()
        else:
            xd = float(x)
            if xd < q[0]:
                q[0] = xd
                k = 1
            elif q[4] <= xd:
                q[4] = xd
                k = 4
            else:
                k = bisect.bisect_right(q, xd)  # std::upper_bound
            for i in range(k, 5):
                pos[i] += 1.0
            for i in range(5):
                des[i] += self.INC[i]
            for i in (1, 2, 3):
                d = des[i] - pos[i]
                dp = pos[i + 1] - pos[i]
                dm = pos[i - 1] - pos[i]
                hp = (q[i + 1] - q[i]) / dp
                hm = (q[i - 1] - q[i]) / dm
                if (d >= 1.0 and dp > 1.0) or (d <= -1.0 and dm < -1.0):
                    s = float(int(d / abs(d)))
                    h = q[i] + s / (dp - dm) * ((s - dm) * hp + (dp - s) * hm)
                    if q[i - 1] < h < q[i + 1]:
                        q[i] = h
                    else:
                        if d > 0:
                            q[i] += hp
                        if d < 0:
                            q[i] -= hm
                    pos[i] += s
        if n > 1:
            mean_n = float(self.S) / float(n)
            tmp = float(x) - mean_n
            self.var = (self.var * float(n - 1)) / float(n) + (tmp * tmp) / float(n - 1)

    def results(self):
        return (
            u16_from_double(float(self.S) / float(self.n)),
            u16_from_double(self.q[2]),
            u16_from_double(self.var),
        )


def build(seqs, function_index, seq_id=None):
    """seqs: list[bytes]; returns (rows in table order, stats dict).

    rows: list of (kmer bytes, avg_from_end, function_index, mean, median, var)
    """
    if seq_id is None:
        seq_id = list(range(len(seqs)))
    table = {}  # kmer -> list of attrs, newest first (TBB multimap order inside a key)
    n_occ = 0
    seqs_with_func = {}
    for s, f, sid in zip(seqs, function_index, seq_id):
        assert f != UNDEFINED
        seqs_with_func[f] = seqs_with_func.get(f, 0) + 1
        L = len(s)
        for p in range(0, L - K + 1):
            kmer = bytes(s[p : p + K])
            if all(c in OK_PROT for c in kmer):
                table.setdefault(kmer, []).insert(0, (f, (L - p) & 0xFFFF, sid, L))
                n_occ += 1
    rows = []
    sig_seqs = set()
    distinct_functions = {}
    for kmer, items in table.items():
        count = len(items)
        func_count = {}
        for it in items:
            func_count[it[0]] = func_count.get(it[0], 0) + 1
        best_func, best_count = UNDEFINED, -1
        for f in sorted(func_count):  # std::map order, strict >
            if best_func == UNDEFINED or func_count[f] > best_count:
                if best_func == UNDEFINED:
                    best_func, best_count = f, func_count[f]
                elif func_count[f] > best_count:
                    best_func, best_count = f, func_count[f]
        thresh = np.float32(count) * np.float32(0.8)
        if np.float32(best_count) < thresh:
            continue
        acc = BoostAcc()
        offsets = []
        for f, off, sid, L in items:
            if f == best_func:
                acc.push(L)
            offsets.append(off)
            sig_seqs.add(sid)
        mean, median, var = acc.results()
        offsets.sort()
        rows.append((kmer, offsets[len(offsets) // 2], best_func, mean, median, var))
        distinct_functions[best_func] = distinct_functions.get(best_func, 0) + 1
    def table_order(r):
        # presentation only: the order include/sigk.h defines for the table (all-upper-case k-mers first, in byte
        # order; then the others by case-folded bytes and the case mask with residue j in bit j)
        k = r[0]
        mask = sum(((c >> 5) & 1) << j for j, c in enumerate(k))
        return (mask != 0, bytes(c & 0xDF for c in k), mask)

    rows.sort(key=table_order)
    stats = dict(
        n_occurrences=n_occ,
        n_distinct_kmers=len(table),
        distinct_signatures=len(rows),
        num_seqs_with_a_signature=len(sig_seqs),
        distinct_functions=distinct_functions,
        seqs_with_func=seqs_with_func,
    )
    return rows, stats
