"""Full-size parity on the GPU box: the GPU kept table of a BASELINE.json configuration against the CPU oracle,
plus size-independent properties of the GPU table.  Used by tests/test_gpu_fullsize.py (pytest -m gpu) and runnable
by hand:

  python tests/fullsize_check.py config2 [n_threads] [n_proteins_prefix]

Tier A (kmer, function_index, avg_from_end, mean, counters) does not depend on the order in which equal k-mers were
inserted, so the threaded oracle checks it on every row; tier B (median, var) is defined in canonical insertion
order (the reference at --n-threads 1), so it is checked against the 1-thread oracle.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle_c  # noqa: E402
from signature_kmers_b200.builder import GpuSignatureBuilder  # noqa: E402
from signature_kmers_b200.capi import kmer_case_masks  # noqa: E402
from signature_kmers_b200.synth import Synth  # noqa: E402


def table_properties(got, n_proteins):
    """Properties of a kept table that need no oracle (any size)."""
    n_up = got.n_upper
    assert 0 <= n_up <= got.n_kept
    masks = kmer_case_masks(got.kmer)
    assert not masks[:n_up].any() and masks[n_up:].all(), "two sections: k-mers without / with a lower-case residue"
    be = got.kmer.view(">u8").ravel().astype(np.uint64)
    assert (be[1:n_up] > be[:n_up - 1]).all() if n_up > 1 else True, "first section strictly increasing in byte order"
    folded = be[n_up:] & np.uint64(0xDFDFDFDFDFDFDFDF)
    if len(folded) > 1:
        m2 = masks[n_up:].astype(np.uint64)
        assert ((folded[1:] > folded[:-1]) | ((folded[1:] == folded[:-1]) & (m2[1:] > m2[:-1]))).all(), "second section increasing"
    assert got.n_kept == got.distinct_signatures == int(got.distinct_functions.sum())
    assert int(got.seqs_with_func.sum()) == n_proteins
    assert (got.function_index != 0xFFFF).all()
    assert got.num_seqs_with_a_signature <= n_proteins


def run(workload, threads, prefix=None, overrides=None, log=print):
    """Build `workload` (or its first `prefix` proteins) on GPU 0 and compare with the oracle on `threads` threads."""
    synth = Synth.config(workload, **(overrides or {})) if workload in ("config1", "config2", "config3", "config4") else Synth(**overrides)
    p = synth.packed(0, prefix) if prefix else synth.packed()
    b = GpuSignatureBuilder(device=0)
    b.set_proteins(p)
    t0 = time.time()
    got = b.build()
    tm = b.timings()
    b.close()
    log(f"[{workload}] gpu build+fetch {time.time() - t0:.1f}s: {p.n_proteins} proteins, {got.n_occurrences} occ, {got.n_distinct_kmers} groups, "
        f"{got.n_kept} kept ({got.n_kept - got.n_upper} with a lower-case residue), device {tm['device_total_ms']:.1f} ms")
    table_properties(got, p.n_proteins)
    t0 = time.time()
    want, secs = oracle_c.oracle_build(p, n_threads=threads)
    log(f"[{workload}] oracle ({threads} threads) extract+process {secs:.1f}s, total {time.time() - t0:.1f}s")
    assert got.n_occurrences == want.n_occurrences and got.n_distinct_kmers == want.n_distinct_kmers and got.n_kept == want.n_kept
    assert got.num_seqs_with_a_signature == want.num_seqs_with_a_signature and got.n_upper == want.n_upper
    for col in ("kmer", "function_index", "avg_from_end", "mean", "distinct_functions", "seqs_with_func"):
        assert np.array_equal(getattr(got, col), getattr(want, col)), col
    if threads == 1:
        for col in ("median", "var"):
            assert np.array_equal(getattr(got, col), getattr(want, col)), col
    else:
        same = (got.median == want.median) & (got.var == want.var)
        log(f"[{workload}] order-dependent columns equal on {100.0 * same.mean():.2f}% of rows (the threaded oracle's order is nondeterministic)")
    tier = "A+B" if threads == 1 else "A"
    log(f"FULLSIZE_CHECK_PASSED {workload}{' prefix %d' % prefix if prefix else ''}: tier {tier} bit-exact on {got.n_kept} rows")
    return got.n_kept


if __name__ == "__main__":
    wl = sys.argv[1] if len(sys.argv) > 1 else "config2"
    th = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
    pre = int(sys.argv[3]) if len(sys.argv) > 3 else None
    run(wl, th, pre)
