"""GPU FASTA parser (csrc/fasta.cu behind sigk_fasta_parse / sigk_fasta_commit) against the host
reader of signature_kmers_b200/host — which tests/test_host_dropin.py pins to the REFERENCE's own FastaParser —
and against a plain restatement of the five-state machine (src/fasta_parser.h:38-144) for the reported characters.
Bit-exact: ids, definitions, sequences, record order, error positions and states."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from signature_kmers_b200 import capi
from tests.test_host_dropin import FASTA_CASES, parse_with
from tests.util import assert_tables_equal, pack, random_proteins

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "signature_kmers_b200")
EMPTY = (b"", b"", b"")


@pytest.fixture(scope="module")
def gpu():
    from signature_kmers_b200.builder import GpuSignatureBuilder

    b = GpuSignatureBuilder(device=0)
    yield b
    b.close()


@pytest.fixture(scope="module")
def host():
    so = os.path.join(PKG, "libsigk_host.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-C", os.path.join(PKG, "host"), "../libsigk_host.so"], check=True, capture_output=True)
    lib = C.CDLL(so)
    lib.sigk_host_fasta_parse.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64]
    lib.sigk_host_fasta_parse.restype = C.c_uint64
    return lib


def machine(data: bytes):
    """FastaParser::parse_char, one byte at a time: (records, errors) with errors = [(position, state)]."""
    START, ID, DEF, DATA, LINE = range(5)
    st, recs, errs = START, [], []
    cur = None
    for pos, c in enumerate(data):
        ch = bytes([c])
        if ch == b"\r":
            continue
        alpha = ch.isalpha() and c < 128
        if st == START:
            if ch == b">":
                cur = [b"", b"", b""]; recs.append(cur); st = ID
            else:
                errs.append((pos, START))
        elif st == ID:
            if ch in b" \t":
                cur[1] += ch; st = DEF
            elif ch == b"\n":
                st = DATA
            else:
                cur[0] += ch
        elif st == DEF:
            if ch == b"\n":
                st = DATA
            else:
                cur[1] += ch
        elif st == DATA:
            if ch == b"\n":
                st = LINE
            elif alpha or ch == b"*":
                cur[2] += ch
            else:
                errs.append((pos, DATA))
        else:
            if ch == b">":
                cur = [b"", b"", b""]; recs.append(cur); st = ID
            elif ch == b"\n":
                pass
            elif alpha:
                cur[2] += ch; st = DATA
            else:
                errs.append((pos, LINE))
    return [tuple(r) for r in recs], errs


def gpu_records(gpu, files):
    """Per file: the records the device parser found, as (id, def, seq), and its reported characters as (offset in the file, state)."""
    p = gpu.fasta_parse(files)
    stream = gpu.dbg_fasta_stream(p["n_residues"]).tobytes()
    buf = p["bytes"].tobytes()
    per_file = [[] for _ in files]
    per_file_err = [[] for _ in files]
    ends = p["file_begin"] + p["file_len"]
    for r in range(p["n_records"]):
        hp = int(p["header_pos"][r])
        f = int(np.searchsorted(p["file_begin"], hp, side="right")) - 1
        assert hp < ends[f]
        end = int(ends[f])
        ie = int(p["id_end"][r]); ie = end if ie == capi.SIGK_FASTA_NO_POS else ie
        le = int(p["line_end"][r]); le = end if le == capi.SIGK_FASTA_NO_POS else le
        assert hp < ie <= le <= end
        rid = buf[hp + 1:ie].replace(b"\r", b"")
        rdef = buf[ie:le].replace(b"\r", b"")
        seq = stream[int(p["seq_begin"][r]):int(p["seq_begin"][r + 1])]
        per_file[f].append((rid, rdef, seq))
    assert p["n_errors"] == len(p["errors"]) or p["n_errors"] > capi.SIGK_FASTA_MAX_ERRORS
    for e in p["errors"]:
        pos, st = int(e) & ((1 << 60) - 1), int(e) >> 60
        f = int(np.searchsorted(p["file_begin"], pos, side="right")) - 1
        per_file_err[f].append((pos - int(p["file_begin"][f]), st))
    return per_file, per_file_err, p


def callback_sequence(records):
    """What the reference's callback sees for one file: the records, then parse()'s parse_complete and the caller's second one."""
    return list(records) + ([EMPTY] if records else [EMPTY, EMPTY])


def test_cases_one_file_each(gpu, host):
    files = list(FASTA_CASES)
    got, errs, _ = gpu_records(gpu, files)
    for data, recs, er in zip(files, got, errs):
        assert callback_sequence(recs) == parse_with(host.sigk_host_fasta_parse, data), data
        want_recs, want_errs = machine(data)
        assert recs == want_recs and er == want_errs, data


def test_fuzz_many_files_in_one_parse(gpu, host):
    rng = np.random.default_rng(321)
    alphabet = np.frombuffer(b">>\n\n\n\r \tACDEFGHIKLMNPQRSTVWYXacx**19-|[]#\x00\xe9", dtype=np.uint8)
    weights = np.ones(len(alphabet))
    weights[8:32] = 6.0
    weights /= weights.sum()
    files = []
    for trial in range(500):
        n = int(rng.integers(0, 700))
        data = bytes(rng.choice(alphabet, size=n, p=weights))
        if trial % 3 == 0:
            data = b">" + data
        files.append(data)
    got, errs, p = gpu_records(gpu, files)
    assert p["n_errors"] < capi.SIGK_FASTA_MAX_ERRORS
    for data, recs, er in zip(files, got, errs):
        assert callback_sequence(recs) == parse_with(host.sigk_host_fasta_parse, data), data
        want_recs, want_errs = machine(data)
        assert recs == want_recs and er == want_errs, data


@pytest.mark.parametrize("size", [8191, 8192, 8193, 16384, 3 * 8192 + 5])
def test_tile_boundaries(gpu, host, size):
    """Headers, CRLF pairs, blank lines and the '>' of the next record placed across the 8 KB tile and the 16-byte
    thread boundaries."""
    rng = np.random.default_rng(size)
    for shift in range(0, 40, 3):
        body = bytearray()
        k = 0
        while len(body) < size + 64:
            k += 1
            body += b">id%d def %d [g]\r\n" % (k, k) if k % 3 else b">h%d\n" % k
            for _ in range(int(rng.integers(0, 4))):
                body += bytes(rng.choice(np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8), size=int(rng.integers(1, 90)))) + (b"\r\n" if k % 2 else b"\n")
            if k % 5 == 0:
                body += b"\n\n*AC\n"
        data = b"x" * shift + bytes(body[:size])
        got, errs, _ = gpu_records(gpu, [data, data[:size // 2], b"", data[7:]])
        for d, recs, er in zip([data, data[:size // 2], b"", data[7:]], got, errs):
            assert callback_sequence(recs) == parse_with(host.sigk_host_fasta_parse, d)
            assert er == machine(d)[1]


def test_long_lines_and_headers(gpu, host):
    big = b">id1 some definition [genome]\n" + b"ACDEFGHIKLMNPQRSTVWY" * 9000 + b"\n>id2\n" + (b"ACDEFGHIKL\n" * 20000) + b">" + b"i" * 70000 + b" d" * 40000 + b"\nAC*DE\n"
    got, errs, p = gpu_records(gpu, [big])
    assert callback_sequence(got[0]) == parse_with(host.sigk_host_fasta_parse, big)
    assert errs[0] == [] and p["n_errors"] == 0
    assert [len(r[2]) for r in got[0]] == [180000, 200000, 5]


def to_fasta(seqs, ids, width=60, crlf=False):
    nl = b"\r\n" if crlf else b"\n"
    out = bytearray()
    for i, s in zip(ids, seqs):
        out += b">" + i + b" hypothetical protein [genome]" + nl
        for k in range(0, len(s), width):
            out += s[k:k + width] + nl
    return bytes(out)


def test_commit_feeds_the_build(gpu):
    """Proteins written as FASTA files, parsed and committed on the device, give the table of the same proteins handed
    over as arrays — with records dropped by the caller (deleted ids, ids without a function) in between."""
    seqs, funcs = random_proteins(41, n_families=60, members=(2, 14), length=(20, 400))
    seqs = [s.replace(b"*", b"X") for s in seqs]         # ('*' at the start of a wrapped line would be reported and dropped, as in the reference)
    ids = [b"fig|%d.peg.%d" % (i % 7, i) for i in range(len(seqs))]
    n_files = 5
    cuts = [len(seqs) * k // n_files for k in range(n_files + 1)]
    files = [to_fasta(seqs[a:b], ids[a:b], width=60 + 7 * k, crlf=bool(k % 2)) for k, (a, b) in enumerate(zip(cuts, cuts[1:]))]
    p = gpu.fasta_parse(files)
    assert p["n_records"] == len(seqs) and p["n_errors"] == 0
    assert np.array_equal(np.diff(p["seq_begin"]), np.array([len(s) for s in seqs], dtype=np.uint64))
    keep = np.ones(len(seqs), dtype=np.uint8)
    keep[::9] = 0                                        # what the caller's id lookups would drop
    seq_id = np.arange(len(seqs), dtype=np.uint32) + 1000
    gpu.fasta_commit(keep, np.asarray(funcs, dtype=np.uint16), seq_id)
    got = gpu.build()
    kept = [i for i in range(len(seqs)) if keep[i]]
    gpu.set_proteins(pack([seqs[i].decode() for i in kept], [funcs[i] for i in kept], seq_id=seq_id[kept]))
    want = gpu.build()
    assert_tables_equal(got, want, tier_b=True)
    assert got.n_kept > 0 and got.num_seqs_with_a_signature == want.num_seqs_with_a_signature


def test_commit_without_a_parse_is_refused():
    from signature_kmers_b200.builder import GpuSignatureBuilder

    b = GpuSignatureBuilder(device=0)
    with pytest.raises(capi.SigkError):
        b.fasta_commit(np.zeros(1, np.uint8), np.zeros(1, np.uint16), np.zeros(1, np.uint32))
    p = b.fasta_parse([])
    assert p["n_records"] == 0 and p["n_residues"] == 0
    b.fasta_commit(np.zeros(0, np.uint8), np.zeros(0, np.uint16), np.zeros(0, np.uint32))
    t = b.build()
    assert t.n_kept == 0
    b.close()


def test_cli_gpu_fasta_equals_host_reader(tmp_path):
    """kmers-build-signatures --gpu-fasta: the k-mer pass reads the FASTA files through the device parser; every
    output file is the one the default run (host reader) writes.  The tree has deleted ids, ids without a function,
    CRLF files and a header-only record, so that the gating around the parser is exercised too."""
    from signature_kmers_b200.synth import Synth

    subprocess.run(["make", "-C", os.path.join(PKG, "host")], check=True, capture_output=True)
    s = Synth(n_proteins=3000, n_functions=80, n_genomes=6, seed=23)
    tree = tmp_path / "tree"
    s.write_tree(str(tree))
    seqs = sorted((tree / "Seqs").iterdir())
    # awkward files: CRLF line ends; a record without sequence in front of another; an id that has no assignment
    text = seqs[0].read_bytes()
    seqs[0].write_bytes(text.replace(b"\n", b"\r\n"))
    first_id = seqs[1].read_bytes().split(b"\n", 1)[0][1:].split()[0]
    seqs[1].write_bytes(b">lonely\n" + seqs[1].read_bytes() + b">nobody.knows.me\nACDEFGHIKLMNPQRSTVWY\n")
    deleted = tmp_path / "deleted.txt"
    deleted.write_bytes(first_id + b"\n")
    outs = []
    for flag in ([], ["--gpu-fasta"]):
        out = tmp_path / ("out" + ("_gpu" if flag else ""))
        cmd = [os.path.join(PKG, "kmers-build-signatures"), "-D", str(tree / "Annotations" / "0"), "-F", str(tree / "Seqs"),
               "--kmer-data-dir", str(out), "--final-kmers", "final.kmers", "--sigk-table", "kmer_data.sigk", "--sorted-files", "--n-threads", "3",
               "--deleted-features-file", str(deleted)] + flag
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        outs.append((out, r.stdout))
    (a, a_out), (b, b_out) = outs
    assert a_out == b_out
    names = sorted(p.relative_to(a).as_posix() for p in a.rglob("*") if p.is_file())
    assert names == sorted(p.relative_to(b).as_posix() for p in b.rglob("*") if p.is_file())
    assert "final.kmers" in names and "kmer_data.sigk" in names
    for n in names:
        assert (a / n).read_bytes() == (b / n).read_bytes(), n
    assert (a / "final.kmers").stat().st_size > 1000
