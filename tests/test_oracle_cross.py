"""The two oracle restatements (C++ faithful-structure, pure Python) agree on
random ragged inputs; the threaded C++ form agrees with the serial one on the
order-independent (tier A) fields."""
import numpy as np
import pytest

from oracle import oracle_py
from tests.util import assert_tables_equal, pack, random_proteins


@pytest.mark.parametrize("seed", range(6))
def test_cpp_vs_python(oracle, seed):
    seqs, funcs = random_proteins(seed, n_families=10, alphabet=b"ACDEFGHIKL" if seed % 2 else b"ACDEFGHIKLMNPQRSTVWY")
    t, _ = oracle.oracle_build(pack(seqs, funcs))
    rows, stats = oracle_py.build(seqs, funcs)
    assert t.n_kept == len(rows)
    assert t.n_occurrences == stats["n_occurrences"]
    assert t.n_distinct_kmers == stats["n_distinct_kmers"]
    assert t.num_seqs_with_a_signature == stats["num_seqs_with_a_signature"]
    got = [(bytes(t.kmer[i]), int(t.avg_from_end[i]), int(t.function_index[i]), int(t.mean[i]), int(t.median[i]), int(t.var[i]))
           for i in range(t.n_kept)]
    assert got == rows
    for f, c in stats["distinct_functions"].items():
        assert int(t.distinct_functions[f]) == c
    for f, c in stats["seqs_with_func"].items():
        assert int(t.seqs_with_func[f]) == c


def test_heavy_duplication_exercises_psquare(oracle):
    # few distinct residues, many members: segments far above 5 samples with mixed lengths
    seqs, funcs = random_proteins(99, n_families=3, members=(40, 60), length=(30, 90), sub_rate=0.02,
                                  ambig_rate=0.0, lower_rate=0.0, alphabet=b"ACDE")
    t, _ = oracle.oracle_build(pack(seqs, funcs))
    rows, _ = oracle_py.build(seqs, funcs)
    got = [(bytes(t.kmer[i]), int(t.avg_from_end[i]), int(t.function_index[i]), int(t.mean[i]), int(t.median[i]), int(t.var[i]))
           for i in range(t.n_kept)]
    assert got == rows
    assert (t.median > 0).any() and (t.var > 0).any()


@pytest.mark.parametrize("threads", [2, 5])
def test_threaded_oracle_tier_a(oracle, threads):
    seqs, funcs = random_proteins(7, n_families=30, members=(1, 12))
    p = pack(seqs, funcs)
    a, _ = oracle.oracle_build(p, n_threads=1)
    b, _ = oracle.oracle_build(p, n_threads=threads)
    assert_tables_equal(a, b, tier_b=False)
