"""Shared helpers for the parity tests: small random protein sets that exercise
shared k-mers, mixed functions, ambiguity codes and ragged lengths."""
from __future__ import annotations

import numpy as np

from signature_kmers_b200.capi import KeptTable, PackedProteins

AA = b"ACDEFGHIKLMNPQRSTVWY"
AMBIG = b"XBZUJO*xbz"


def random_proteins(seed: int, n_families: int = 12, members=(1, 9), length=(5, 80), sub_rate=0.1,
                    ambig_rate=0.01, lower_rate=0.01, n_functions=None, alphabet=AA, shared_domain=True):
    rng = np.random.default_rng(seed)
    n_functions = n_functions or n_families
    seqs, funcs = [], []
    alpha = np.frombuffer(alphabet, dtype=np.uint8)
    domain = alpha[rng.integers(0, len(alpha), 16)]
    for f in range(n_families):
        L = int(rng.integers(length[0], length[1] + 1))
        anc = alpha[rng.integers(0, len(alpha), L)].copy()
        if shared_domain and f % 3 == 0 and L > 20:
            anc[2:18] = domain
        for _ in range(int(rng.integers(members[0], members[1] + 1))):
            s = anc.copy()
            mut = rng.random(L) < sub_rate
            s[mut] = alpha[rng.integers(0, len(alpha), int(mut.sum()))]
            amb = rng.random(L) < ambig_rate
            s[amb] = np.frombuffer(AMBIG, dtype=np.uint8)[rng.integers(0, len(AMBIG), int(amb.sum()))]
            low = rng.random(L) < lower_rate
            s[low] = s[low] | 0x20
            if L > 0 and rng.random() < 0.3:  # indel: drop a run of residues (ragged lengths inside a family)
                cut = int(rng.integers(0, L))
                s = np.delete(s, slice(cut, cut + int(rng.integers(1, 12))))
            seqs.append(s.tobytes())
            funcs.append(f % n_functions)
    order = rng.permutation(len(seqs))
    seqs = [seqs[i] for i in order]
    funcs = [funcs[i] for i in order]
    return seqs, funcs


def pack(seqs, funcs, seq_id=None) -> PackedProteins:
    return PackedProteins.from_sequences(seqs, funcs, seq_id)


def assert_tables_equal(a: KeptTable, b: KeptTable, tier_b: bool = True, what: str = ""):
    assert a.n_occurrences == b.n_occurrences, what
    assert a.n_distinct_kmers == b.n_distinct_kmers, what
    assert a.n_kept == b.n_kept, what
    assert a.distinct_signatures == b.distinct_signatures, what
    assert a.num_seqs_with_a_signature == b.num_seqs_with_a_signature, what
    np.testing.assert_array_equal(a.kmer, b.kmer, err_msg=what)
    np.testing.assert_array_equal(a.function_index, b.function_index, err_msg=what)
    np.testing.assert_array_equal(a.avg_from_end, b.avg_from_end, err_msg=what)
    np.testing.assert_array_equal(a.mean, b.mean, err_msg=what)
    np.testing.assert_array_equal(a.distinct_functions, b.distinct_functions, err_msg=what)
    np.testing.assert_array_equal(a.seqs_with_func, b.seqs_with_func, err_msg=what)
    if tier_b:
        np.testing.assert_array_equal(a.median, b.median, err_msg=what)
        np.testing.assert_array_equal(a.var, b.var, err_msg=what)


def reorder_to_table_order(kmers, cols):
    """(list of k-mer strings, list of per-row arrays) in any order -> the same rows in the table order of include/sigk.h."""
    from signature_kmers_b200.capi import table_order_key

    order = sorted(range(len(kmers)), key=lambda i: table_order_key(kmers[i]))
    idx = np.array(order, dtype=np.int64)
    return [kmers[i] for i in order], [np.asarray(c)[idx] if len(order) else np.asarray(c) for c in cols]
