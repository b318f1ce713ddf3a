"""Generates the golden fixtures of this directory FROM THE REFERENCE'S OWN SOURCES.

Run in the build container, where /root/reference exists:

    make -C oracle ref && python tests/golden/make_golden.py

`oracle/_ref/libref_signature.so` and `libref_call.so` are the reference's signature builder and function caller
(`/root/reference/src/signature_build.{h,tcc}`, `function_map.h`, `seed_utils.h`, `call_functions.{h,tcc}`, ...)
compiled unmodified over the stand-in Boost/TBB headers of `oracle/refshim/` (see the README there for what that does
and does not pin).  For every case this script writes the input tree and what the reference code produced from it:

    <case>/tree/Annotations/0/*, <case>/tree/Seqs/*     inputs (plus good_*.txt / ignored.txt / deleted.txt if any)
    <case>/table.tsv.gz     kmer, avg_from_end, function_index, mean, median, var   (rows sorted by k-mer bytes)
    <case>/counters.json    Kept / distinct_signatures / num_seqs_with_a_signature, distinct_functions, seqs_with_func
    <case>/function.index   as the reference's writer printed it
    <case>/queries.fa, <case>/calls.txt                 call side: FASTA queries and the reference caller's output
                                                        ("#call" region lines, then id, function, index, score)
    <case>/expected/stdout.txt, distinct_functions, recall.report.d/*
                                                        what the reference's whole command line (its main, compiled
                                                        as oracle/_ref/ref-kmers-build-signatures) printed and wrote

tests/test_golden.py checks the CPU oracle, the drop-in's host code and (on a GPU) the whole drop-in against these
files; nothing there needs /root/reference or oracle/_ref.
"""
import ctypes as C
import gzip
import json
import os
import random
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from signature_kmers_b200.synth import Synth  # noqa: E402

AA = "ACDEFGHIKLMNPQRSTVWY"


def load_refs():
    sig = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_signature.so"))
    sig.ref_signature_build_ex.argtypes = [C.c_char_p] * 6 + [C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_ulonglong), C.POINTER(C.c_uint),
                                                               C.POINTER(C.c_uint)]
    call = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_call.so"))
    u16p = np.ctypeslib.ndpointer(dtype=np.uint16, flags="C_CONTIGUOUS")
    call.ref_call_functions.argtypes = [C.c_ulonglong, C.c_char_p, u16p, u16p, u16p, u16p, u16p, C.c_char_p, C.c_char_p, C.c_ulonglong,
                                        C.c_int, C.c_int, C.c_char_p, C.c_ulonglong]
    call.ref_call_functions.restype = C.c_ulonglong
    return sig, call


def read_table(path):
    raw = open(path, "rb").read()
    n = int(np.frombuffer(raw, dtype=np.uint64, count=1, offset=8)[0])
    kmers = raw[16:16 + 8 * n]
    off = 16 + 8 * n
    cols = [np.frombuffer(raw, dtype=np.uint16, count=n, offset=off + 2 * n * c).copy() for c in range(5)]
    return n, kmers, cols


def finish_case(sig, call, case_dir, queries, opts):
    tree = os.path.join(case_dir, "tree")
    out = os.path.join(case_dir, "_out")
    os.makedirs(out, exist_ok=True)
    counters = (C.c_ulonglong * 3)()
    df = (C.c_uint * 65536)()
    swf = (C.c_uint * 65536)()
    arg = lambda name: os.path.join(tree, name).encode() if opts.get(name) else b""
    rc = sig.ref_signature_build_ex(os.path.join(tree, "Annotations", "0").encode(), os.path.join(tree, "Seqs").encode(), arg("deleted.txt"),
                                    arg("good_functions.txt"), arg("good_roles.txt"), arg("ignored.txt"), 3, 1, out.encode(), counters, df, swf)
    assert rc == 0
    n, kmers, cols = read_table(os.path.join(out, "ref_table.bin"))
    text = "".join("%s\t%d\t%d\t%d\t%d\t%d\n" % (kmers[8 * i:8 * i + 8].decode("latin-1"), cols[0][i], cols[1][i], cols[2][i], cols[3][i], cols[4][i])
                   for i in range(n))
    with open(os.path.join(case_dir, "table.tsv.gz"), "wb") as raw, gzip.GzipFile(filename="", fileobj=raw, mode="wb", mtime=0) as f:
        f.write(text.encode("latin-1"))          # mtime 0: regenerating gives identical bytes
    json.dump({"kept": int(counters[0]), "distinct_signatures": int(counters[1]), "num_seqs_with_a_signature": int(counters[2]),
               "distinct_functions": {str(i): int(v) for i, v in enumerate(df) if v},
               "seqs_with_func": {str(i): int(v) for i, v in enumerate(swf) if v}},
              open(os.path.join(case_dir, "counters.json"), "w"), indent=0, sort_keys=True)
    shutil.copy(os.path.join(out, "function.index"), os.path.join(case_dir, "function.index"))
    fasta = "".join(">%s\n%s\n" % q for q in queries)
    open(os.path.join(case_dir, "queries.fa"), "w").write(fasta)
    buf = C.create_string_buffer(1 << 22)
    data = fasta.encode("latin-1")
    m = call.ref_call_functions(n, kmers, *cols, os.path.join(out, "function.index").encode(), data, len(data), 0, 1, buf, len(buf))
    assert m <= len(buf)
    open(os.path.join(case_dir, "calls.txt"), "wb").write(buf.raw[:m])
    shutil.rmtree(out)
    # the reference's whole command line on the same tree: stdout and the files it derives from the table
    main_out = os.path.join(case_dir, "_main")
    cmd = [os.path.join(ROOT, "oracle", "_ref", "ref-kmers-build-signatures"), "-D", os.path.join(tree, "Annotations", "0"),
           "-F", os.path.join(tree, "Seqs"), "--kmer-data-dir", main_out, "--final-kmers", "final.kmers", "--min-reps-required", "3",
           "--n-threads", "1"]
    for flag, name in (("--good-functions", "good_functions.txt"), ("--good-roles", "good_roles.txt"),
                       ("--ignored-functions-file", "ignored.txt"), ("--deleted-features-file", "deleted.txt")):
        if opts.get(name):
            cmd += [flag, os.path.join(tree, name)]
    import subprocess
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=case_dir)
    assert r.returncode == 0, r.stderr
    exp = os.path.join(case_dir, "expected")
    shutil.rmtree(exp, ignore_errors=True)
    os.makedirs(os.path.join(exp, "recall.report.d"))
    # stdout without the three path lines (they echo the caller's directories)
    open(os.path.join(exp, "stdout.txt"), "w").write("".join(l + "\n" for l in r.stdout.splitlines()[3:]))
    lines = sorted(open(os.path.join(main_out, "distinct_functions")).read().splitlines(), key=lambda l: int(l.split("\t")[0]))
    open(os.path.join(exp, "distinct_functions"), "w").write("".join(l + "\n" for l in lines))        # index order (the reference: hash order)
    for f in os.listdir(os.path.join(main_out, "recall.report.d")):
        shutil.copy(os.path.join(main_out, "recall.report.d", f), os.path.join(exp, "recall.report.d", f))
    # final.kmers is the first three columns of table.tsv.gz; checked here, not stored twice
    want = sorted(open(os.path.join(main_out, "final.kmers")).read().splitlines())
    assert want == ["%s\t%d\t%d\t" % (kmers[8 * i:8 * i + 8].decode("latin-1"), cols[0][i], cols[1][i]) for i in range(n)]
    # (the main reads the files in readdir order, the fixtures above were made in sorted order: the statistics columns of
    # function.index and median / var depend on that order, names, counters, calls and recall reports do not)
    names = lambda path: [l.split("\t")[:2] for l in open(path).read().splitlines()]
    assert names(os.path.join(main_out, "function.index")) == names(os.path.join(case_dir, "function.index"))
    shutil.rmtree(main_out)
    print("%s: %d kept k-mers, %d queries" % (os.path.basename(case_dir), n, len(queries)))


def records_of_tree(tree):
    recs = []
    for f in sorted(os.listdir(os.path.join(tree, "Seqs"))):
        for blk in open(os.path.join(tree, "Seqs", f)).read().split(">")[1:]:
            head, _, body = blk.partition("\n")
            recs.append((head.split()[0], body.replace("\n", "")))
    return recs


def queries_from(recs, seed, n_train, n_chimera):
    rng = random.Random(seed)
    q = [r for r in recs[:: max(1, len(recs) // n_train)]][:n_train]
    for i in range(n_chimera):
        a, b = rng.choice(recs)[1], rng.choice(recs)[1]
        if len(a) < 20 or len(b) < 20:
            continue
        s = a[:rng.randrange(10, len(a))] + b[rng.randrange(0, len(b) - 9):]
        s = "".join(c if rng.random() > 0.03 else rng.choice(AA + "X*") for c in s)
        q.append(("chimera%d" % i, s))
    return q


def case_synthetic(sig, call):
    d = os.path.join(HERE, "synthetic")
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    Synth(n_proteins=150, n_functions=12, n_genomes=4, seed=71).write_tree(os.path.join(d, "tree"))
    expected = os.path.join(d, "tree", "function.index.expected")
    if os.path.exists(expected):
        os.remove(expected)
    recs = records_of_tree(os.path.join(d, "tree"))
    finish_case(sig, call, d, queries_from(recs, 1, 25, 15), {})


def case_zipf(sig, call):
    """Skewed family sizes and little mutation: k-mers shared by dozens of proteins, so the P-square markers move,
    the 16-bit length sum wraps and the 80 % rule sees mixed groups."""
    d = os.path.join(HERE, "zipf")
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    Synth(n_proteins=260, n_functions=10, n_genomes=5, seed=72, zipf_s=1.1, mut_rate=0.02).write_tree(os.path.join(d, "tree"))
    expected = os.path.join(d, "tree", "function.index.expected")
    if os.path.exists(expected):
        os.remove(expected)
    recs = records_of_tree(os.path.join(d, "tree"))
    finish_case(sig, call, d, queries_from(recs, 3, 25, 15), {})


def case_edge(sig, call):
    """Hand-made: the 80 % rule at its boundary, ambiguity codes and case, short proteins, missing / un-kept / ignored
    functions, comments and truncation markers, definition-line functions, good roles and functions, a deleted
    feature, a wrapping 16-bit length sum."""
    d = os.path.join(HERE, "edge")
    shutil.rmtree(d, ignore_errors=True)
    tree = os.path.join(d, "tree")
    os.makedirs(os.path.join(tree, "Annotations", "0"))
    os.makedirs(os.path.join(tree, "Seqs"))
    core = "MKTAYIAKQRQISFVKSHFSRQLEERLGLIEV"
    other = "GSHMLEDPVAGTWQNCYRFKAGDTLSKIAEEH"

    def seq(tag, n):
        return "".join(AA[(i * 7 + tag * 3 + (i // 5)) % 20] for i in range(n))

    for gi in range(4):
        g = "3000%d.1" % gi
        ann, fa = [], []

        def add(fn, s, definition=""):
            rid = "fig|%s.peg.%d" % (g, len(fa) + 1)
            if fn is not None:
                ann.append((rid, fn))
            fa.append((rid, definition, s))

        add("Alpha synthase (EC 1.1.1.1) # a note", core + "ACDEFGHIKLMNPQRS" * (gi + 1))
        add("Alpha synthase (EC 1.1.1.1)", core[:20] + "WWWWWWWWWW")
        add("Beta kinase", other + core[:12])
        add("Beta kinase", other[:10] + "X" + other[11:] + "acdefghikl")
        add("Beta kinase # truncated", seq(2, 70))
        add("Gamma lyase", "ACDEFGH")
        add(None, core)
        add("Rare thing %d" % gi, core[5:25])
        add("Delta ligase", ("QWERTYIPASDFGHKLCVNM" * 4)[: 60 + gi])
        add("hypothetical protein", other + "TTTTTTTTTTTT")
        add(None, seq(8, 85), " Iota reductase")
        add("Kappa oxidase", seq(9, 85), " Something else entirely")
        add("Ignored function", seq(10, 85))
        if gi < 2:
            add("Epsilon pump @ Zeta channel", seq(4, 100))
            add("Listed function", seq(6, 100))
        for rep in range(60):
            add("Eta pump", "HHHHHHHHKKKKKKKK" + "ACDEFGHIKLMNPQRSTVWY" * (10 + (rep * 7 + gi) % 9))
        with open(os.path.join(tree, "Annotations", "0", g), "w") as f:
            for rid, fn in ann:
                f.write("%s\t%s\n" % (rid, fn))
        with open(os.path.join(tree, "Seqs", g), "w") as f:
            for rid, definition, s in fa:
                f.write(">%s%s\n%s\n" % (rid, definition, s))
    for gi in range(3):
        with open(os.path.join(tree, "Seqs", "genbank%d" % gi), "w") as f:
            f.write(">prot%d_a Lambda transferase [Some organism %d]\n%s\n" % (gi, gi, seq(12, 95)))
            f.write(">prot%d_b Lambda transferase # truncated [Some organism %d]\n%s\n" % (gi, gi, seq(12, 60)))
    open(os.path.join(tree, "good_functions.txt"), "w").write("Listed function\n")
    open(os.path.join(tree, "good_roles.txt"), "w").write("Zeta channel\n")
    open(os.path.join(tree, "ignored.txt"), "w").write("Ignored function\n")
    open(os.path.join(tree, "deleted.txt"), "w").write("fig|30001.1.peg.3\n")
    recs = records_of_tree(tree)
    finish_case(sig, call, d, queries_from(recs, 2, 30, 15),
                {"good_functions.txt": 1, "good_roles.txt": 1, "ignored.txt": 1, "deleted.txt": 1})


if __name__ == "__main__":
    sig, call = load_refs()
    case_synthetic(sig, call)
    case_zipf(sig, call)
    case_edge(sig, call)
