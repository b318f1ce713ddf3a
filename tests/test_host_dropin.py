"""Host side of the drop-in (signature_kmers_b200/host): FASTA reader against
the REFERENCE's own parser (oracle/_ref, compiled from /root/reference in place
when present), the hand-matched SEED regexes, and the command line's gating /
packing against the synthetic generator."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from signature_kmers_b200.synth import Synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "signature_kmers_b200")


@pytest.fixture(scope="module")
def host():
    subprocess.run(["make", "-C", os.path.join(PKG, "host"), "../libsigk_host.so"], check=True, capture_output=True)
    lib = C.CDLL(os.path.join(PKG, "libsigk_host.so"))
    for name in ("sigk_host_fasta_parse",):
        getattr(lib, name).argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64]
        getattr(lib, name).restype = C.c_uint64
    for name in ("sigk_host_split_func_comment", "sigk_host_roles"):
        getattr(lib, name).argtypes = [C.c_char_p, C.c_char_p, C.c_uint64]
        getattr(lib, name).restype = C.c_uint64
    lib.sigk_host_is_truncated.argtypes = [C.c_char_p]
    return lib


def parse_with(fn, data: bytes):
    out = C.create_string_buffer(4 * len(data) + 64)
    n = fn(data, len(data), out, len(out))
    recs = out.raw[:n].split(b"\x02")[:-1]
    return [tuple(r.split(b"\x01")) for r in recs]


FASTA_CASES = [
    b">a\nACDE\nFGH\n>b desc here\nKLMN\n",
    b">a\r\nACDE\r\nFG*H\r\n>b\tdef [g]\r\nKL\r\n",                # CRLF, mid-line '*', tab before def
    b">a\nAC\n\n\nDE\n>b\n*KL\nMN\n",                               # blank lines, '*' at line start is dropped
    b">hdronly\n>next def\nACDE\n",                                 # header-only record swallows the next header
    b">a\nACDE",                                                    # no final newline
    b"junk\n>a\nAC1DE\n>b\nxyz\n",                                  # garbage before '>', digit in data, lower case
    b"",
    b">a  two blanks\nAC DE\n",
]


@pytest.mark.parametrize("data", FASTA_CASES)
def test_fasta_reader_matches_reference_parser(host, data):
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libref_fasta.so")
    if not os.path.exists(ref_so):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True, capture_output=True)
    if not os.path.exists(ref_so):
        pytest.skip("oracle/_ref not built (reference checkout absent)")
    ref = C.CDLL(ref_so)
    ref.ref_fasta_parse.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64]
    ref.ref_fasta_parse.restype = C.c_uint64
    assert parse_with(host.sigk_host_fasta_parse, data) == parse_with(ref.ref_fasta_parse, data)


def test_fasta_reader_known_records(host):
    # golden records (generated with the reference parser, see the test above)
    assert parse_with(host.sigk_host_fasta_parse, FASTA_CASES[0]) == [(b"a", b"", b"ACDEFGH"), (b"b", b" desc here", b"KLMN"), (b"", b"", b"")]
    assert parse_with(host.sigk_host_fasta_parse, FASTA_CASES[3])[0] == (b"hdronly", b"", b"nextdefACDE")


@pytest.mark.parametrize("s,want", [
    ("Alpha beta", ("Alpha beta", "", "")),
    ("Alpha beta # fragment", ("Alpha beta", "#", "fragment")),
    ("Alpha  ##  note # more", ("Alpha", "##", "note # more")),
    ("Alpha #nospace", ("Alpha #nospace", "", "")),
    ("Alpha ##x # real", ("Alpha ##x", "#", "real")),
    ("  # lead", ("", "#", "lead")),
    ("", ("", "", "")),
])
def test_split_func_comment(host, s, want):
    out = C.create_string_buffer(1024)
    host.sigk_host_split_func_comment(s.encode(), out, len(out))
    assert tuple(out.value.decode().split("\x01")) == want


def test_truncation_and_roles(host):
    assert host.sigk_host_is_truncated(b"fragment") and host.sigk_host_is_truncated(b"truncated") and host.sigk_host_is_truncated(b"missing x")
    assert not host.sigk_host_is_truncated(b"a fragment")
    out = C.create_string_buffer(1024)
    host.sigk_host_roles(b"Role A / Role B @ Role C; Role D # comment", out, len(out))
    assert out.value.decode().split("\x01")[:-1] == ["Role A", "Role B", "Role C", "Role D"]
    host.sigk_host_roles(b"A/B;C", out, len(out))                  # no blanks: not delimiters
    assert out.value.decode().split("\x01")[:-1] == ["A/B;C"]


@pytest.mark.parametrize("n,threads", [(0, 1), (1, 4), (5, 2), (1 << 20, 3), ((1 << 21) + 17, 4)])
def test_final_kmers_writer(host, tmp_path, n, threads):
    """KMER \\t avg_from_end \\t function_index \\t \\n (src/kmers-build-signatures.cc:213-217), blocks formatted in parallel."""
    rng = np.random.default_rng(n + 1)
    kmers = rng.choice(np.frombuffer(b"ACDEFGHIKLMNPQRSTVWYacd", dtype=np.uint8), size=(n, 8))
    avg = rng.integers(0, 65536, size=n).astype(np.uint16)
    fn = rng.integers(0, 65536, size=n).astype(np.uint16)
    if n:
        avg[0], fn[0] = 0, 65535
    u16p = np.ctypeslib.ndpointer(dtype=np.uint16, flags="C_CONTIGUOUS")
    host.sigk_host_write_final_kmers.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, u16p, u16p, C.c_int]
    path = str(tmp_path / "final.kmers")
    assert host.sigk_host_write_final_kmers(path.encode(), n, kmers.tobytes(), avg, fn, threads) == 0
    got = open(path, "rb").read()
    if n <= 5:
        want = b"".join(bytes(k) + b"\t%d\t%d\t\n" % (a, f) for k, a, f in zip(kmers, avg, fn))
        assert got == want
    else:
        lines = got.split(b"\n")
        assert lines[-1] == b"" and len(lines) == n + 1
        for i in [i for i in list(range(0, n, 104729)) + [n - 1, (1 << 20) - 1, 1 << 20] if i < n]:
            assert lines[i] == bytes(kmers[i]) + b"\t%d\t%d\t" % (avg[i], fn[i]), i


def read_packed(path):
    raw = open(path, "rb").read()
    np_, total = np.frombuffer(raw, dtype=np.uint64, count=2)
    np_, total = int(np_), int(total)
    off = 16
    starts = np.frombuffer(raw, dtype=np.uint64, count=np_ + 1, offset=off); off += 8 * (np_ + 1)
    func = np.frombuffer(raw, dtype=np.uint16, count=np_, offset=off); off += 2 * np_
    sid = np.frombuffer(raw, dtype=np.uint32, count=np_, offset=off); off += 4 * np_
    res = np.frombuffer(raw, dtype=np.uint8, count=total, offset=off)
    return res, starts, func, sid


def test_cli_gating_and_packing_match_generator(tmp_path):
    """Generator tree -> drop-in command line (--dump-packed, no GPU) == generator's packed arrays:
    FunctionMap (genome evidence, >= 3 genomes, std::set order of indices), seq_id arithmetic, gates."""
    if not os.path.exists(os.path.join(PKG, "libsigk.so")):
        import __graft_entry__ as g
        g.build()
    subprocess.run(["make", "-C", os.path.join(PKG, "host")], check=True, capture_output=True)
    # 60 families of 5..6 members over 4 genomes + 6 families of 2 members (never reach 3 genomes: gated out)
    s = Synth(n_proteins=330, n_functions=66, n_genomes=4, seed=11, zipf_s=0.0)
    tree = tmp_path / "tree"
    s.write_tree(str(tree))
    out = tmp_path / "out"
    dump = tmp_path / "packed.bin"
    r = subprocess.run([os.path.join(PKG, "kmers-build-signatures"), "-D", str(tree / "Annotations" / "0"), "-F", str(tree / "Seqs"),
                        "--kmer-data-dir", str(out), "--final-kmers", "final.kmers", "--sorted-files", "--dump-packed", str(dump)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    res, starts, func, sid = read_packed(dump)
    p = s.packed()
    np.testing.assert_array_equal(starts, p.starts)
    np.testing.assert_array_equal(res, p.residues)
    np.testing.assert_array_equal(func, p.function_index)
    np.testing.assert_array_equal(sid, p.seq_id)
    # function.index: idx \t name ... in std::set order, "hypothetical protein" included
    want = [l.split("\t")[:2] for l in open(tree / "function.index.expected").read().splitlines()]
    got = [l.split("\t")[:2] for l in open(out / "function.index").read().splitlines()]
    assert got == want and ["hypothetical protein"] in [g[1:] for g in got]
    assert f"kept {len(want)} functions" in r.stdout
    assert open(out / "genomes").read() == "empty genomes\n" and open(out / "otu.index").read() == ""


@pytest.mark.gpu
def test_cli_end_to_end_final_kmers(tmp_path, oracle):
    """The drop-in command line on a synthetic tree: final.kmers (k-mer, avg_from_end, function_index) equals the oracle."""
    from signature_kmers_b200.capi import PackedProteins

    subprocess.run(["make", "-C", os.path.join(PKG, "host")], check=True, capture_output=True)
    s = Synth(n_proteins=3000, n_functions=150, n_genomes=6, seed=12)
    tree = tmp_path / "tree"
    s.write_tree(str(tree))
    out = tmp_path / "out"
    cmd = [os.path.join(PKG, "kmers-build-signatures"), "-D", str(tree / "Annotations" / "0"), "-F", str(tree / "Seqs"),
           "--kmer-data-dir", str(out), "--final-kmers", "final.kmers", "--min-reps-required", "3", "--n-threads", "4"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    dump = tmp_path / "packed.bin"
    subprocess.run(cmd + ["--dump-packed", str(dump)], check=True, capture_output=True)   # same readdir order
    res, starts, func, sid = read_packed(dump)
    want, _ = oracle.oracle_build(PackedProteins(res.copy(), starts.copy(), func.copy(), sid.copy()))
    rows = [l.split("\t") for l in open(out / "final.kmers").read().splitlines()]
    assert len(rows) == want.n_kept and all(len(x) == 4 and x[3] == "" for x in rows)       # trailing tab
    assert [x[0] for x in rows] == want.kmer_strings()
    assert [int(x[1]) for x in rows] == [int(v) for v in want.avg_from_end]
    assert [int(x[2]) for x in rows] == [int(v) for v in want.function_index]
    assert f"Kept {want.n_kept} kmers" in r.stdout
    assert f"distinct_signatures={want.distinct_signatures}" in r.stdout
    assert f"num_seqs_with_a_signature={want.num_seqs_with_a_signature}" in r.stdout
    df = {int(l.split("\t")[0]): int(l.split("\t")[2]) for l in open(out / "distinct_functions").read().splitlines()}
    assert df == {i: int(c) for i, c in enumerate(want.distinct_functions) if c}
    assert os.path.isdir(out / "recall.report.d")


def test_fasta_reader_fuzz_against_reference_parser(host):
    """Random byte soup over the characters the state machine distinguishes, in many chunkings of the input
    (the reader appends runs of ordinary characters in bulk; the reference goes one character at a time)."""
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libref_fasta.so")
    if not os.path.exists(ref_so):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], capture_output=True)
    if not os.path.exists(ref_so):
        pytest.skip("oracle/_ref not built (reference checkout absent)")
    ref = C.CDLL(ref_so)
    ref.ref_fasta_parse.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64]
    ref.ref_fasta_parse.restype = C.c_uint64
    rng = np.random.default_rng(123)
    alphabet = np.frombuffer(b">>\n\n\n\r \tACDEFGHIKLMNPQRSTVWYXacx**19-|[]#", dtype=np.uint8)
    weights = np.ones(len(alphabet))
    weights[8:32] = 6.0                                              # mostly letters, so that records form
    weights /= weights.sum()
    for trial in range(400):
        n = int(rng.integers(0, 400))
        data = bytes(rng.choice(alphabet, size=n, p=weights))
        if trial % 3 == 0:
            data = b">" + data
        assert parse_with(host.sigk_host_fasta_parse, data) == parse_with(ref.ref_fasta_parse, data), data
    # long sequences and headers crossing the reader's 64 KB chunks
    big = b">id1 some definition [genome]\n" + b"ACDEFGHIKLMNPQRSTVWY" * 9000 + b"\n>id2\n" + (b"ACDEFGHIKL\n" * 20000) + b">" + b"i" * 70000 + b" d" * 40000 + b"\nAC*DE\n"
    assert parse_with(host.sigk_host_fasta_parse, big) == parse_with(ref.ref_fasta_parse, big)
