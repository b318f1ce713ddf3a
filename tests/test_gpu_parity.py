"""GPU parity tests: every stage of the CUDA path, through the C ABI, against
the CPU oracle / numpy on the same seeded inputs (bit-exact: integer work).
Run on the B200 box with `pytest -m gpu`."""
import numpy as np
import pytest

from signature_kmers_b200 import capi
from tests.util import assert_tables_equal, pack, random_proteins

pytestmark = pytest.mark.gpu

ALPHABET = "ACDEFGHIKLMNPQRSTVWYacdefghiklmnpqrstvwy"
SYM = {ord(c): i for i, c in enumerate(ALPHABET)}       # rank = SYM % 20, lower case = SYM >= 20


@pytest.fixture(scope="module")
def gpu():
    from signature_kmers_b200.builder import GpuSignatureBuilder

    b = GpuSignatureBuilder(device=0)
    yield b
    b.close()


def encode_reference(seqs):
    """Window enumeration of src/signature_build.tcc:162-180 in pure Python."""
    codes, ords, offs = [], [], []
    for i, s in enumerate(seqs):
        L = len(s)
        for p in range(L - 7):
            w = s[p:p + 8]
            if all(c in SYM for c in w):
                # group code: base-20 code of the case-folded residues << 8 | case mask (residue j = bit j)
                code, mask = 0, 0
                for j, c in enumerate(w):
                    code = code * 20 + SYM[c] % 20
                    mask |= (SYM[c] >= 20) << j
                codes.append((code << 8) | mask)
                ords.append(i)
                offs.append((L - p) & 0xFFFF)
    return np.array(codes, dtype=np.uint64), np.array(ords, dtype=np.uint32), np.array(offs, dtype=np.uint16)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_encode_matches_window_loop(gpu, seed):
    seqs, funcs = random_proteins(seed, n_families=40, members=(1, 8), length=(0, 300), ambig_rate=0.02, lower_rate=0.03)
    seqs = seqs + [b"", b"ACDEFGH", b"ACDEFGHI", b"X" * 40, b"acdefghiklmnpqrstvwy" * 300]  # empty, short, exact, all-bad, > 4096 tile
    funcs = funcs + [0] * 5
    code, ordinal, offset = gpu.dbg_encode(pack(seqs, funcs))
    rc, ro, rf = encode_reference(seqs)
    np.testing.assert_array_equal(code, rc)
    np.testing.assert_array_equal(ordinal, ro)
    np.testing.assert_array_equal(offset, rf)


def test_encode_long_protein_offset_wraps(gpu):
    L = 70000
    rng = np.random.default_rng(5)
    s = bytes(rng.choice(np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8), L))
    code, ordinal, offset = gpu.dbg_encode(pack([b"ACDEFGHIK", s, b"ACDEFGHIK"], [0, 1, 2]))
    rc, ro, rf = encode_reference([b"ACDEFGHIK", s, b"ACDEFGHIK"])
    np.testing.assert_array_equal(code, rc)
    np.testing.assert_array_equal(ordinal, ro)
    np.testing.assert_array_equal(offset, rf)


@pytest.mark.parametrize("n,lo,hi", [(1, 21, 64), (3328, 21, 64), (3329, 21, 64), (7681, 21, 64), (100_000, 0, 64), (5_000_000, 21, 64), (1_000_003, 5, 13), (300_000, 30, 39)])
def test_radix_sort_is_stable_and_sorted(gpu, n, lo, hi):
    rng = np.random.default_rng(n)
    # heavy duplication in the sorted bits so that stability is visible
    pool = rng.integers(0, 1 << 62, size=max(1, n // 7), dtype=np.uint64)
    keys = pool[rng.integers(0, len(pool), size=n)]
    keys[: n // 10] = keys[0]                      # one huge run (skewed digit)
    vals = np.arange(n, dtype=np.uint32)
    gk, gv = gpu.dbg_sort_pairs(keys, vals, lo, hi)
    field = (keys >> np.uint64(lo)) & np.uint64((1 << (hi - lo)) - 1) if hi - lo < 64 else keys
    order = np.argsort(field, kind="stable")
    np.testing.assert_array_equal(gk, keys[order])
    np.testing.assert_array_equal(gv, vals[order])


def test_known_answers_on_gpu(gpu, oracle):
    aa = "ACDEFGHIKLMNPQRSTVWY"
    gpu.set_proteins(pack([aa] * 3, [0, 0, 0]))
    t = gpu.build()
    assert t.n_occurrences == 39 and t.n_kept == 13 and t.num_seqs_with_a_signature == 3
    assert t.row("ACDEFGHI") == dict(avg_from_end=20, function_index=0, mean=20, median=20, var=0)
    # K4: the 16-bit sum wraps
    s2 = "WWWWWWWW" + "A" * 292
    gpu.set_proteins(pack([s2] * 300, [0] * 300))
    t = gpu.build()
    assert t.row("WWWWWWWW")["mean"] == 81 and t.row("WWWWWWWW")["avg_from_end"] == 300
    # K2: 4 of 5 kept, 3 of 4 rejected
    gpu.set_proteins(pack([aa[:8]] * 5, [0, 0, 0, 0, 1]))
    assert gpu.build().n_kept == 1
    gpu.set_proteins(pack([aa[:8]] * 4, [0, 0, 0, 1]))
    assert gpu.build().n_kept == 0


def test_empty_and_degenerate_inputs(gpu, oracle):
    for seqs, funcs in ([], []), ([b""], [0]), ([b"ACDEFGH"], [3]), ([b"XXXXXXXXXXXX"], [1]):
        p = pack(seqs, funcs)
        gpu.set_proteins(p)
        got = gpu.build()
        want, _ = oracle.oracle_build(p)
        assert_tables_equal(got, want, what=f"{seqs}")
        assert got.n_kept == 0


@pytest.mark.parametrize("seed,kw", [
    (11, dict(n_families=30, members=(1, 9), length=(5, 120))),
    (12, dict(n_families=8, members=(20, 60), length=(30, 200), sub_rate=0.02, alphabet=b"ACDE")),     # big groups, P^2 active
    (13, dict(n_families=200, members=(1, 20), length=(8, 400), sub_rate=0.15, n_functions=17)),         # mixed functions
    (14, dict(n_families=3, members=(300, 400), length=(300, 330), sub_rate=0.01, ambig_rate=0.0)),      # sums wrap, groups > 255
])
def test_full_build_matches_oracle(gpu, oracle, seed, kw):
    seqs, funcs = random_proteins(seed, **kw)
    p = pack(seqs, funcs)
    gpu.set_proteins(p)
    got = gpu.build()
    want, _ = oracle.oracle_build(p)
    assert_tables_equal(got, want, tier_b=True, what=f"seed {seed}")
    # rows are in table order (include/sigk.h) and unique
    k = got.kmer_strings()
    assert k == sorted(k, key=capi.table_order_key) and len(set(k)) == len(k)
    assert got.n_upper == sum(1 for x in k if not any(c.islower() for c in x))


def test_seq_id_collisions_and_gaps(gpu, oracle):
    seqs, funcs = random_proteins(21, n_families=20, members=(1, 6), length=(20, 90))
    sid = (np.arange(len(seqs)) // 2 * 1000).astype(np.uint32)       # pairs share an id, ids are sparse
    p = pack(seqs, funcs, sid)
    gpu.set_proteins(p)
    got = gpu.build()
    want, _ = oracle.oracle_build(p)
    assert_tables_equal(got, want)


def test_medium_build_many_tiles(gpu, oracle):
    # ~3 M occurrences: several hundred sort tiles, look-back chains across waves
    seqs, funcs = random_proteins(31, n_families=400, members=(10, 40), length=(200, 500), sub_rate=0.12, n_functions=300)
    p = pack(seqs, funcs)
    gpu.set_proteins(p)
    got = gpu.build()
    want, _ = oracle.oracle_build(p, n_threads=1)
    assert got.n_occurrences > 2_000_000
    assert_tables_equal(got, want, tier_b=True)
    # idempotence: a second build on the same handle gives the same table
    again = gpu.build()
    assert_tables_equal(got, again)


def test_split_calls_equal_one_call(gpu):
    seqs, funcs = random_proteins(41, n_families=50, members=(2, 10), length=(50, 150))
    p = pack(seqs, funcs)
    gpu.set_proteins(p)
    a = gpu.build()
    gpu.upload(); gpu.build_device(); gpu.build_device(); gpu.download()
    assert_tables_equal(a, gpu.result())


def test_inlined_division_is_ieee(gpu):
    """csrc/length_acc.cuh divides with its own inlined sequence (the order statistics of a giant group are one
    dependent chain of divisions): it must give the bits of __ddiv_rn, which are the bits of the host's division
    (the reference's Boost accumulators run on the host FPU, round to nearest)."""
    rng = np.random.default_rng(11)
    n = 1 << 20
    parts_a, parts_b = [], []
    # the operand shapes of the recurrences: height differences over small integers, var * (n-1) over n, +-1 over integers
    parts_a.append(rng.integers(-70000, 70000, n).astype(np.float64) + rng.random(n)); parts_b.append(rng.integers(1, 400000, n).astype(np.float64))
    parts_a.append(rng.random(n) * 1e9); parts_b.append(-rng.integers(1, 400000, n).astype(np.float64))
    parts_a.append(np.where(rng.random(n) < 0.5, 1.0, -1.0)); parts_b.append(rng.integers(2, 1 << 22, n).astype(np.float64))
    # arbitrary bit patterns (all exponents, denormals, infinities, NaN) and exact zeros
    parts_a.append(rng.integers(0, 1 << 64, n, dtype=np.uint64).view(np.float64)); parts_b.append(rng.integers(0, 1 << 64, n, dtype=np.uint64).view(np.float64))
    parts_a.append(np.array([0.0, -0.0, 0.0, -0.0, 5e-324, 1e-310, 1.7e308, 1e-300, 1.0, np.inf, 3.0, np.nan]))
    parts_b.append(np.array([3.0, 3.0, -3.0, -3.0, 3.0, 7.0, 1e-10, 1e300, 0.0, 2.0, np.inf, 2.0]))
    a, b = np.concatenate(parts_a), np.concatenate(parts_b)
    inl, lib = gpu.dbg_ddiv(a, b)
    with np.errstate(all="ignore"):
        want = a / b
    nan = np.isnan(want)
    assert np.array_equal(np.isnan(inl), nan) and np.array_equal(np.isnan(lib), nan)
    assert np.array_equal(inl.view(np.uint64)[~nan], lib.view(np.uint64)[~nan])
    assert np.array_equal(inl.view(np.uint64)[~nan], want.view(np.uint64)[~nan])


def test_long_and_giant_groups(gpu, oracle):
    """Groups of 33 .. >1024 records (whole-warp walks, the sampled giant pre-pass),
    mixed functions inside long groups, lengths that differ so P^2 and the
    variance recurrence run for thousands of samples."""
    seqs, funcs = [], []
    # one homopolymer per size: a run of L identical residues gives L-7 windows of one k-mer
    for letter, k in zip("WYFHKMNQ", (33, 64, 512, 513, 1023, 1024, 1025, 3000)):
        seqs.append(letter * (k + 7)); funcs.append(1)
    # a giant shared by many proteins of different lengths, 90 % one function
    rng = np.random.default_rng(3)
    for i in range(60):
        seqs.append("A" * int(rng.integers(200, 2500))); funcs.append(0 if i % 10 else 2)
    # a long group that is rejected (60/40)
    for i in range(50):
        seqs.append("C" * 40 + "DEFGHIKL"); funcs.append(3 if i % 5 < 3 else 4)
    # a giant one (> 4096 records: the CTA-wide reduce) that is rejected, and a giant one kept with equal lengths
    for i in range(12):
        seqs.append("G" * 700); funcs.append(6 if i % 5 < 3 else 7)
    for i in range(9):
        seqs.append("P" * 600); funcs.append(8)
    # filler so that groups straddle chunk and tile boundaries at odd places
    fs, ff = random_proteins(77, n_families=40, members=(2, 12), length=(30, 200))
    seqs += [x.decode() for x in fs]; funcs += [5 + f for f in ff]
    p = pack(seqs, funcs)
    gpu.set_proteins(p)
    got = gpu.build()
    want, _ = oracle.oracle_build(p)
    assert_tables_equal(got, want, tier_b=True)
    assert got.row("AAAAAAAA") is not None and got.row("CCCCCCCC") is None
    assert got.row("GGGGGGGG") is None and got.row("PPPPPPPP")["median"] == 600
    assert got.row("WWWWWWWW")["avg_from_end"] == want.row("WWWWWWWW")["avg_from_end"]


@pytest.mark.parametrize("form", ["compact", "wide"])
def test_both_meta_table_forms(oracle, monkeypatch, form):
    """The per-protein table is 4 or 8 bytes wide depending on job size; both forms, forced, on the same inputs
    (groups of every kind: single records, packed groups, whole-warp groups, ordered walks)."""
    from signature_kmers_b200.builder import GpuSignatureBuilder

    monkeypatch.setenv("SIGK_TEST_META", form)
    seqs, funcs = random_proteins(91, n_families=50, members=(2, 40), length=(30, 400), sub_rate=0.05, n_functions=30)
    seqs = [bytes(x) for x in seqs] + [b"W" * 600, b"W" * 500, b"W" * 77, b"Y" * 40000]
    funcs = list(funcs) + [1, 1, 2, 3]
    p = pack(seqs, funcs)
    b = GpuSignatureBuilder(device=0)
    b.set_proteins(p)
    got = b.build()
    want, _ = oracle.oracle_build(p)
    assert_tables_equal(got, want, tier_b=True, what=form)
    b.close()


def test_proteins_longer_than_65535(gpu, oracle):
    """A protein of 65 535 residues or more switches the per-protein table from the compact 4-byte form to the
    8-byte one; its length enters sums (mod 65536), the equal-length shortcut and the P^2 / variance walks in full."""
    rng = np.random.default_rng(8)
    aa = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)
    big = bytes(rng.choice(aa, 66_000))
    seqs = [big, big[:65_535], big[100:66_000], big[:70] + bytes(rng.choice(aa, 300))]
    funcs = [0, 0, 0, 0]
    fs, ff = random_proteins(78, n_families=20, members=(3, 9), length=(40, 300))
    seqs += [big[:50] + bytes(x) for x in fs[:6]]          # short proteins sharing k-mers with the long ones
    funcs += [0, 0, 0, 1, 0, 0]
    seqs += list(fs)
    funcs += [2 + f for f in ff]
    p = pack(seqs, funcs)
    gpu.set_proteins(p)
    got = gpu.build()
    want, _ = oracle.oracle_build(p)
    assert_tables_equal(got, want, tier_b=True)
    # and right below the threshold the compact form is used: same answers as the oracle either way
    seqs2 = [big[:65_534]] + seqs[3:]
    p2 = pack(seqs2, funcs[:1] + funcs[3:])
    gpu.set_proteins(p2)
    got2 = gpu.build()
    want2, _ = oracle.oracle_build(p2)
    assert_tables_equal(got2, want2, tier_b=True)


@pytest.mark.parametrize("seed", [51, 52, 53])
def test_boundary_alignment_sweep(gpu, oracle, seed):
    """Many small inputs whose group boundaries fall on every alignment of the
    32-record windows, 256-record chunks and 2048-record tiles."""
    rng = np.random.default_rng(seed)
    for _ in range(6):
        seqs, funcs = random_proteins(int(rng.integers(1 << 30)), n_families=int(rng.integers(1, 30)), members=(1, 40),
                                      length=(8, 120), sub_rate=float(rng.choice([0.0, 0.02, 0.2])), alphabet=b"ACDEFG",
                                      n_functions=int(rng.integers(1, 6)))
        p = pack(seqs, funcs)
        gpu.set_proteins(p)
        got = gpu.build()
        want, _ = oracle.oracle_build(p)
        assert_tables_equal(got, want, tier_b=True)


def test_no_order_stats_flag(oracle):
    from signature_kmers_b200 import capi
    from signature_kmers_b200.builder import GpuSignatureBuilder

    seqs, funcs = random_proteins(61, n_families=20, members=(2, 30), length=(30, 200), sub_rate=0.05)
    p = pack(seqs, funcs)
    b = GpuSignatureBuilder(device=0, flags=capi.SIGK_F_NO_ORDER_STATS)
    b.set_proteins(p)
    got = b.build()
    want, _ = oracle.oracle_build(p)
    assert_tables_equal(got, want, tier_b=False)
    assert not got.median.any() and not got.var.any()
    b.close()
