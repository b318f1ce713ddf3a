"""Golden fixtures generated from the reference's own sources (tests/golden/make_golden.py): the CPU oracle, the
drop-in's host code and — on a GPU — the whole drop-in must reproduce them.  Nothing here needs /root/reference or
oracle/_ref: the fixtures are committed."""
import ctypes as C
import gzip
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle_c
from signature_kmers_b200.capi import PackedProteins
from tests.test_host_dropin import read_packed
from tests.util import reorder_to_table_order

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "signature_kmers_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = ["synthetic", "zipf", "edge"]


def golden_table(case):
    rows = [l.rstrip("\n").split("\t") for l in gzip.open(os.path.join(GOLDEN, case, "table.tsv.gz"), "rt")]
    kmers = [r[0] for r in rows]
    cols = [np.array([int(r[c]) for r in rows], dtype=np.uint16) for c in range(1, 6)]
    kmers, cols = reorder_to_table_order(kmers, cols)      # the fixture lists rows in byte order
    return kmers, cols, json.load(open(os.path.join(GOLDEN, case, "counters.json")))


def cli_args(case, out):
    tree = os.path.join(GOLDEN, case, "tree")
    args = [os.path.join(PKG, "kmers-build-signatures"), "-D", os.path.join(tree, "Annotations", "0"), "-F", os.path.join(tree, "Seqs"),
            "--kmer-data-dir", str(out), "--sorted-files"]
    for flag, name in (("--good-functions", "good_functions.txt"), ("--good-roles", "good_roles.txt"),
                       ("--ignored-functions-file", "ignored.txt"), ("--deleted-features-file", "deleted.txt")):
        if os.path.exists(os.path.join(tree, name)):
            args += [flag, os.path.join(tree, name)]
    return args


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(os.path.join(PKG, "libsigk.so")):
        import __graft_entry__ as g
        g.build()
    subprocess.run(["make", "-C", os.path.join(PKG, "host")], check=True, capture_output=True)


def check_table(table_kmers, table_cols, counters_got, case):
    kmers, cols, counters = golden_table(case)
    assert table_kmers == kmers
    for name, a, b in zip(("avg_from_end", "function_index", "mean", "median", "var"), table_cols, cols):
        np.testing.assert_array_equal(a, b, err_msg="%s: %s" % (case, name))
    for key in ("kept", "distinct_signatures", "num_seqs_with_a_signature"):
        assert counters_got[key] == counters[key], key
    assert counters_got["distinct_functions"] == counters["distinct_functions"]
    assert counters_got["seqs_with_func"] == counters["seqs_with_func"]


@pytest.mark.parametrize("case", CASES)
def test_oracle_and_host_code_reproduce_the_reference_outputs(built, tmp_path, case):
    """Drop-in host code (FunctionMap, gates, packing) + CPU oracle == what the reference's sources produced."""
    out = tmp_path / "out"
    dump = str(tmp_path / "packed.bin")
    r = subprocess.run(cli_args(case, out) + ["--dump-packed", dump], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    res, starts, func, sid = read_packed(dump)
    t, _ = oracle_c.oracle_build(PackedProteins(res.copy(), starts.copy(), func.copy(), sid.copy()))
    got = {"kept": t.n_kept, "distinct_signatures": t.distinct_signatures, "num_seqs_with_a_signature": t.num_seqs_with_a_signature,
           "distinct_functions": {str(i): int(v) for i, v in enumerate(t.distinct_functions) if v},
           "seqs_with_func": {str(i): int(v) for i, v in enumerate(t.seqs_with_func) if v}}
    check_table(t.kmer_strings(), [t.avg_from_end, t.function_index, t.mean, t.median, t.var], got, case)
    assert open(out / "function.index").read() == open(os.path.join(GOLDEN, case, "function.index")).read()


@pytest.mark.parametrize("case", CASES)
def test_function_caller_reproduces_the_reference_calls(built, case):
    """host/function_caller.h on the golden table and queries == the reference caller's region calls and best calls."""
    from tests.test_function_caller import load_host

    host = load_host()
    kmers, cols, _ = golden_table(case)
    names = {}
    for line in open(os.path.join(GOLDEN, case, "function.index")).read().splitlines():
        idx, name = line.split("\t")[:2]
        names[int(idx)] = name
    names = [names.get(i, "") for i in range(max(names) + 1)]
    fasta = open(os.path.join(GOLDEN, case, "queries.fa")).read().encode("latin-1")
    out = C.create_string_buffer(1 << 22)
    n = host.sigk_host_call_functions(len(kmers), "".join(kmers).encode("latin-1"), *[np.ascontiguousarray(c) for c in cols],
                                      "\n".join(names).encode(), fasta, len(fasta), 0, 1, out, len(out))
    assert n <= len(out)
    want = open(os.path.join(GOLDEN, case, "calls.txt"), "rb").read()
    assert out.raw[:n] == want
    assert want.count(b"#call") > 5


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_dropin_reproduces_the_reference_outputs(built, tmp_path, case):
    """kmers-build-signatures (host code + GPU build) on the golden tree: table file, final.kmers, counters."""
    out = tmp_path / "out"
    r = subprocess.run(cli_args(case, out) + ["--final-kmers", "final.kmers", "--sigk-table", "kmer_data.sigk", "--no-recall"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = open(out / "kmer_data.sigk", "rb").read()
    n = int(np.frombuffer(raw, dtype=np.uint64, count=1, offset=8)[0])
    kmers = [raw[16 + 8 * i: 24 + 8 * i].decode("latin-1") for i in range(n)]
    cols = [np.frombuffer(raw, dtype=np.uint16, count=n, offset=16 + 8 * n + 2 * n * c) for c in range(5)]
    gk, gcols, counters = golden_table(case)
    assert kmers == gk
    for name, a, b in zip(("avg_from_end", "function_index", "mean", "median", "var"), cols, gcols):
        np.testing.assert_array_equal(a, b, err_msg=name)
    assert "Kept %d kmers" % counters["kept"] in r.stdout
    assert "distinct_signatures=%d" % counters["distinct_signatures"] in r.stdout
    assert "num_seqs_with_a_signature=%d" % counters["num_seqs_with_a_signature"] in r.stdout
    df = {l.split("\t")[0]: int(l.split("\t")[2]) for l in open(out / "distinct_functions").read().splitlines()}
    assert df == counters["distinct_functions"]
    lines = open(out / "final.kmers").read().splitlines()
    assert lines == ["%s\t%d\t%d\t" % (k, a, f) for k, a, f in zip(gk, gcols[0], gcols[1])]
    assert open(out / "function.index").read() == open(os.path.join(GOLDEN, case, "function.index")).read()


@pytest.mark.parametrize("case", CASES)
def test_host_outputs_reproduce_the_reference_command_line(built, tmp_path, case):
    """What the reference's whole command line printed and wrote for the golden trees (expected/), against the files
    the drop-in's host code writes from the golden table: counters, distinct_functions, recall.report.d."""
    tree = os.path.join(GOLDEN, case, "tree")
    kmers, cols, counters = golden_table(case)
    lib = C.CDLL(os.path.join(PKG, "libsigk_host.so"))
    u16p = np.ctypeslib.ndpointer(dtype=np.uint16, flags="C_CONTIGUOUS")
    u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
    lib.sigk_host_outputs.argtypes = [C.c_char_p] * 6 + [C.c_int, C.c_int, C.c_char_p, C.c_uint64, C.c_char_p, u16p, u16p, u16p, u16p, u16p, u32p]
    opt = lambda name: (os.path.join(tree, name) if os.path.exists(os.path.join(tree, name)) else "").encode()
    df = np.zeros(65536, dtype=np.uint32)
    for k, v in counters["distinct_functions"].items():
        df[int(k)] = v
    out = tmp_path / "out"
    rc = lib.sigk_host_outputs(os.path.join(tree, "Annotations", "0").encode(), os.path.join(tree, "Seqs").encode(), opt("good_functions.txt"),
                               opt("good_roles.txt"), opt("ignored.txt"), opt("deleted.txt"), 3, 2, str(out).encode(), len(kmers),
                               "".join(kmers).encode("latin-1"), *[np.ascontiguousarray(c) for c in cols], df)
    assert rc == 0
    exp = os.path.join(GOLDEN, case, "expected")
    assert open(out / "distinct_functions").read() == open(os.path.join(exp, "distinct_functions")).read()
    want = {f: open(os.path.join(exp, "recall.report.d", f)).read() for f in os.listdir(os.path.join(exp, "recall.report.d"))}
    got = {f: open(out / "recall.report.d" / f).read() for f in os.listdir(out / "recall.report.d")}
    assert got == want
    stdout = open(os.path.join(exp, "stdout.txt")).read().splitlines()
    assert stdout[1:4] == ["Kept %d kmers" % counters["kept"], "distinct_signatures=%d" % counters["distinct_signatures"],
                           "num_seqs_with_a_signature=%d" % counters["num_seqs_with_a_signature"]]
    assert open(out / "final.kmers").read().splitlines() == ["%s\t%d\t%d\t" % (k, a, f) for k, a, f in zip(kmers, cols[0], cols[1])]
