"""CPU restatement of the reference's function caller (TEST INFRASTRUCTURE — only tests/ may import it).

Follows /root/reference/src/call_functions.tcc:
  for_each_kmer        src/kmer_data.h:76-102
  HitSet.process       src/call_functions.tcc:34-101
  process_aa_seq       src/call_functions.tcc:262-343
  find_best_call       src/call_functions.tcc:352-659
  statistics           Boost.Math univariate statistics (mean / median / median_absolute_deviation on float
                       vectors, :49-51) and Boost.Accumulators mean of floats (:471-533)

Boost is not in this image and the reference has no tests for this path: parity is UNPINNED against a
reference run.  This file is written independently of the C++ product code (signature_kmers_b200/host/
function_caller.h) — pure Python with numpy.float32 scalars for every float operation — and the two are
compared call by call in tests/test_function_caller.py, next to hand-derived known answers.
"""
from __future__ import annotations

import re

import numpy as np

F32 = np.float32
UNDEFINED = 0xFFFF
K = 8


def for_each_kmer(seq: str):
    """Yield (kmer, offset).  A window is skipped if it holds '*'/'X' or ends right before one (kend >= next_ambig)."""
    n = len(seq)
    ambig = [i for i, c in enumerate(seq) if c in "*X"]
    ai = 0
    p = 0
    while p <= n - K:
        while ai < len(ambig) and ambig[ai] < p:
            ai += 1
        nxt = ambig[ai] if ai < len(ambig) else None
        if nxt is not None and p + K >= nxt:
            p = nxt + 1
            continue
        yield seq[p:p + K], p
        p += 1


# ---- Boost.Math statistics on float vectors ---------------------------------------------------
def bm_mean(values):
    """Four interleaved running means, then a weighted combination (single_pass.hpp, random access + real)."""
    v = [F32(x) for x in values]
    n = len(v)
    mu = [F32(0)] * 4
    i = F32(1)
    body = n - n % 4
    k = 0
    while k < body:
        inv = F32(1) / i
        for j in range(4):
            t = v[k + j] - mu[j]
            t = t * inv
            mu[j] = mu[j] + t
        i = i + F32(1)
        k += 4
    num1 = F32(body) / F32(4)
    num2 = num1 + F32(n % 4)
    while k < n:
        mu[3] = mu[3] + (v[k] - mu[3]) / i
        i = i + F32(1)
        k += 1
    return (num1 * ((mu[0] + mu[1]) + mu[2]) + num2 * mu[3]) / F32(n)


def bm_median(values):
    s = sorted(F32(x) for x in values)
    n = len(s)
    if n % 2 == 0:
        return (s[n // 2 - 1] + s[n // 2]) / F32(2)
    return s[n // 2]


def bm_mad(values):
    c = bm_median(values)
    return bm_median([abs(F32(x) - c) for x in values])


# ---- the caller ---------------------------------------------------------------------------------
class FunctionCaller:
    def __init__(self, table: dict, function_index: list, min_hits=5, max_gap=200, ignore_hypothetical=False):
        """table: kmer (str) -> (avg_from_end, function_index, mean, median, var)."""
        self.table = table
        self.function_index = function_index
        self.min_hits = min_hits
        self.max_gap = max_gap
        self.ignore_hypothetical = ignore_hypothetical
        self.hypo = function_index.index("hypothetical protein")        # the reference exits when it is missing

    def fn(self, idx):
        return "" if idx == UNDEFINED else self.function_index[idx]

    def _process(self, hits, seqlen, current, calls):
        mine = [h for h in hits if h[0][1] == current]
        lengths = [F32(h[0][2]) for h in mine]
        mean_length = bm_mean(lengths)
        median_length = bm_median(lengths)
        mad = bm_mad(lengths)
        if mad == 0:
            mad = F32(30)
        lo = float(mean_length) - 2.0 * float(mad)
        hi = float(mean_length) + 2.0 * float(mad)
        if len(mine) >= self.min_hits and not (seqlen < lo or seqlen > hi):
            calls.append(dict(start=hits[0][1], end=mine[-1][1] + K - 1, count=len(mine), function_index=current,
                              median=int(median_length), mad=mad))
        if hits[-2][0][1] != current and hits[-2][0][1] == hits[-1][0][1]:
            return hits[-2][0][1], hits[-2:]
        return current, []

    def process_aa_seq(self, seq: str):
        calls = []
        hits = []                       # (kdata, pos)
        current = UNDEFINED
        seqlen = float(len(seq))
        for kmer, offset in for_each_kmer(seq):
            kd = self.table.get(kmer)
            if kd is None:
                continue
            if self.ignore_hypothetical and kd[1] == self.hypo:
                continue
            if hits and hits[-1][1] + self.max_gap < offset:
                if len(hits) >= self.min_hits:
                    current, hits = self._process(hits, seqlen, current, calls)
                else:
                    hits = []
            if not hits:
                current = kd[1]
            hits.append((kd, offset))
            if len(hits) > 1 and current != kd[1] and hits[-2][0][1] == hits[-1][0][1]:
                current, hits = self._process(hits, seqlen, current, calls)
        if len(hits) >= self.min_hits:
            current, hits = self._process(hits, seqlen, current, calls)
        return calls

    FUSION_RE = re.compile(r"^W?A[A|W]*W[B|W]*BW?")

    def find_best_call(self, calls):
        """-> (function_index, function, score).  score is a numpy float32."""
        if not calls:
            return UNDEFINED, "", F32(0)
        collapsed = []
        for c in calls:
            if collapsed and collapsed[-1]["function_index"] == c["function_index"]:
                collapsed[-1]["end"] = c["end"]
                collapsed[-1]["count"] += c["count"]
            else:
                collapsed.append(dict(c))
        merged = []
        i = 0
        while i < len(collapsed):
            cur = dict(collapsed[i])
            i += 1
            while (i + 1 < len(collapsed) and cur["function_index"] == collapsed[i + 1]["function_index"]
                   and collapsed[i]["count"] < 5 and cur["count"] + collapsed[i + 1]["count"] >= 10):
                cur["end"] = collapsed[i + 1]["end"]
                cur["count"] += collapsed[i + 1]["count"]
                i += 2
            merged.append(cur)

        if len(merged) > 1:
            func_key, fusion_key = {}, {}
            next_func, next_fusion = ord("A"), ord("W")
            info, sums, counts = {}, {}, {}
            exp = ""
            total = 0
            for c in merged:
                total += c["count"]
                func = self.fn(c["function_index"])
                parts = func.split(" / ")
                fk = ""
                for p in parts:
                    if p not in func_key:
                        func_key[p] = chr(next_func)
                        next_func += 1
                    fk += func_key[p]
                if len(parts) > 1:
                    if fk not in fusion_key:
                        fusion_key[fk] = chr(next_fusion)
                        next_fusion += 1
                    key = fusion_key[fk]
                else:
                    key = func_key[func]
                exp += key
                sums[key] = sums.get(key, F32(0)) + F32(c["median"])
                counts[key] = counts.get(key, 0) + 1
                info[key] = (c["function_index"], func)
            m = self.FUSION_RE.match(exp)
            if m and m.end() == len(exp):
                with np.errstate(all="ignore"):
                    mean = lambda k: sums.get(k, F32(0)) / F32(counts.get(k, 0))
                    a, w, b = mean("A"), mean("W"), mean("B")
                    diff = (a + b) - w
                    frac = abs(diff) / w
                if frac < 0.1:
                    return info["W"][0], info["W"][1], F32(total)

        by_func = {}
        for c in merged:
            by_func[c["function_index"]] = by_func.get(c["function_index"], 0) + c["count"]
        vec = sorted(by_func.items())                  # std::map order: ascending function index
        if len(vec) > 1:
            libstdcxx_partial_sort(vec, 2, lambda a, b: a[1] > b[1])
        offset = vec[0][1] if len(vec) == 1 else vec[0][1] - vec[1][1]
        if offset >= 5:
            return vec[0][0], self.fn(vec[0][0]), F32(vec[0][1])
        if len(vec) >= 2:
            f1, f2 = self.fn(vec[0][0]), self.fn(vec[1][0])
            if f2 > f1:
                f1, f2 = f2, f1
            if len(vec) == 2 or vec[1][1] - vec[2][1] > 2:
                return UNDEFINED, f1 + " ?? " + f2, F32(vec[0][1])
        return UNDEFINED, "", F32(0)

    def call(self, seq: str):
        return self.find_best_call(self.process_aa_seq(seq))


# ---- std::partial_sort as libstdc++ implements it (bits/stl_heap.h, bits/stl_algo.h) ---------------------------
# find_best_call sorts only the first two entries (src/call_functions.tcc:594-599) and then reads vec[2]: what is
# left there is whatever the heap selection happened to leave, so the algorithm is restated step by step.
def _push_heap(v, hole, top, value, comp):
    parent = (hole - 1) // 2
    while hole > top and comp(v[parent], value):
        v[hole] = v[parent]
        hole = parent
        parent = (hole - 1) // 2
    v[hole] = value


def _adjust_heap(v, hole, length, value, comp):
    top = hole
    child = hole
    while child < (length - 1) // 2:
        child = 2 * (child + 1)
        if comp(v[child], v[child - 1]):
            child -= 1
        v[hole] = v[child]
        hole = child
    if length % 2 == 0 and child == (length - 2) // 2:
        child = 2 * (child + 1)
        v[hole] = v[child - 1]
        hole = child - 1
    _push_heap(v, hole, top, value, comp)


def libstdcxx_partial_sort(v, middle, comp):
    """std::partial_sort(v.begin(), v.begin() + middle, v.end(), comp), in place."""
    n = len(v)
    if middle >= 2:                                     # __make_heap(first, middle)
        parent = (middle - 2) // 2
        while True:
            _adjust_heap(v, parent, middle, v[parent], comp)
            if parent == 0:
                break
            parent -= 1
    for i in range(middle, n):                          # __heap_select
        if comp(v[i], v[0]):
            value = v[i]                                # __pop_heap(first, middle, i)
            v[i] = v[0]
            _adjust_heap(v, 0, middle, value, comp)
    last = middle                                       # __sort_heap(first, middle)
    while last > 1:
        last -= 1
        value = v[last]
        v[last] = v[0]
        _adjust_heap(v, 0, last, value, comp)


def format_score(x) -> str:
    """C++ `ostream << float` with the default precision (6 significant digits, %g)."""
    return "%g" % float(x)
