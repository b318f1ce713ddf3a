"""Pins the CPU oracle (and the drop-in's host gating) against the REFERENCE'S OWN signature builder:
oracle/_ref/libref_signature.so is /root/reference/src/signature_build.{h,tcc} + function_map.h + seed_utils.h +
fasta_parser.cc compiled unmodified over stand-in Boost/TBB headers (oracle/refshim/README.md says what that does and
does not pin).  Inputs are directory trees, as the reference's command line takes them; the oracle sees the packed
proteins the drop-in command line derives from the same tree.  The library is prebuilt here (make -C oracle ref) and
travels to the GPU box; without it these tests skip."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle_c
from signature_kmers_b200.capi import PackedProteins
from signature_kmers_b200.synth import Synth
from tests.test_host_dropin import read_packed
from tests.util import reorder_to_table_order

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "signature_kmers_b200")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_signature.so")


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_SO):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], capture_output=True)
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libref_signature.so not built (reference checkout absent)")
    if not os.path.exists(os.path.join(PKG, "libsigk.so")):
        import __graft_entry__ as g
        g.build()
    subprocess.run(["make", "-C", os.path.join(PKG, "host")], check=True, capture_output=True)
    lib = C.CDLL(REF_SO)
    lib.ref_signature_build.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_char_p,
                                        C.POINTER(C.c_ulonglong), C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
    return lib


def read_table(path):
    raw = open(path, "rb").read()
    assert raw[:8] == b"SIGKTBL1"
    n = int(np.frombuffer(raw, dtype=np.uint64, count=1, offset=8)[0])
    kmers = [raw[16 + 8 * i: 24 + 8 * i].decode("latin-1") for i in range(n)]
    off = 16 + 8 * n
    cols = [np.frombuffer(raw, dtype=np.uint16, count=n, offset=off + 2 * n * c) for c in range(5)]
    return reorder_to_table_order(kmers, cols)      # the reference shim writes rows in byte order


def run_reference(ref, tree, out, deleted="", min_reps=3):
    os.makedirs(out, exist_ok=True)
    counters = (C.c_ulonglong * 3)()
    df = (C.c_uint * 65536)()
    swf = (C.c_uint * 65536)()
    rc = ref.ref_signature_build(os.path.join(tree, "Annotations", "0").encode(), os.path.join(tree, "Seqs").encode(), deleted.encode(),
                                 min_reps, 1, str(out).encode(), counters, df, swf)
    assert rc == 0
    kmers, cols = read_table(os.path.join(out, "ref_table.bin"))
    return kmers, cols, list(counters), np.array(df), np.array(swf)


def run_oracle_through_dropin(tree, out, tmp, deleted="", min_reps=3):
    """The drop-in command line does the host side (FunctionMap, gates, packing); the oracle does the build."""
    dump = os.path.join(tmp, "packed.bin")
    cmd = [os.path.join(PKG, "kmers-build-signatures"), "-D", os.path.join(tree, "Annotations", "0"), "-F", os.path.join(tree, "Seqs"),
           "--kmer-data-dir", str(out), "--min-reps-required", str(min_reps), "--sorted-files", "--dump-packed", dump]
    if deleted:
        cmd += ["--deleted-features-file", deleted]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    res, starts, func, sid = read_packed(dump)
    table, _ = oracle_c.oracle_build(PackedProteins(res.copy(), starts.copy(), func.copy(), sid.copy()))
    return table


def assert_same(refres, table, what):
    kmers, cols, counters, df, swf = refres
    assert table.kmer_strings() == kmers, what
    for name, c in zip(("avg_from_end", "function_index", "mean", "median", "var"), cols):
        np.testing.assert_array_equal(getattr(table, name), c, err_msg="%s: %s" % (what, name))
    assert counters == [table.n_kept, table.distinct_signatures, table.num_seqs_with_a_signature], what
    np.testing.assert_array_equal(table.distinct_functions, df, err_msg=what)
    np.testing.assert_array_equal(table.seqs_with_func, swf, err_msg=what)


def function_index_names(path):
    return [l.split("\t")[:2] for l in open(path).read().splitlines()]


@pytest.mark.parametrize("seed,n_proteins,n_functions,genomes,zipf", [(61, 1500, 60, 5, 0.0), (62, 1200, 40, 4, 1.1), (63, 400, 90, 3, 0.0)])
def test_oracle_equals_reference_sources_on_synthetic_trees(ref, tmp_path, seed, n_proteins, n_functions, genomes, zipf):
    s = Synth(n_proteins=n_proteins, n_functions=n_functions, n_genomes=genomes, seed=seed, zipf_s=zipf)
    tree = str(tmp_path / "tree")
    s.write_tree(tree)
    refres = run_reference(ref, tree, str(tmp_path / "ref_out"))
    table = run_oracle_through_dropin(tree, tmp_path / "our_out", str(tmp_path))
    assert len(refres[0]) > 1000
    assert_same(refres, table, "synthetic seed %d" % seed)
    # function.index, whole lines: index, name, count, mean, median, var, dev as the reference's writer prints them
    assert open(tmp_path / "ref_out" / "function.index").read() == open(tmp_path / "our_out" / "function.index").read()


def write_tree(root, genomes):
    """genomes: {name: [(id, function or None, sequence)]} -> Annotations/0/<name> + Seqs/<name>"""
    os.makedirs(os.path.join(root, "Annotations", "0"))
    os.makedirs(os.path.join(root, "Seqs"))
    for g, recs in genomes.items():
        with open(os.path.join(root, "Annotations", "0", g), "w") as a, open(os.path.join(root, "Seqs", g), "w") as f:
            for rid, fn, seq in recs:
                if fn is not None:
                    a.write("%s\t%s\n" % (rid, fn))
                f.write(">%s\n%s\n" % (rid, seq))


def test_oracle_equals_reference_sources_on_edge_cases(ref, tmp_path):
    """Hand-made tree: the 80 % rule at its boundary, ties, ambiguity codes and case, proteins shorter than K,
    proteins without / with un-kept functions (seq_id arithmetic), a deleted feature, comments in assignments,
    a 16-bit length-sum wrap and groups long enough for the P^2 markers to move."""
    core = "MKTAYIAKQRQISFVKSHFSRQLEERLGLIEV"
    other = "GSHMLEDPVAGTWQNCYRFKAGDTLSKIAEEH"
    genomes = {}
    for gi in range(4):
        g = "1000%d.1" % gi
        recs = []
        n = 0

        def add(fn, seq):
            nonlocal n
            n += 1
            recs.append(("fig|%s.peg.%d" % (g, n), fn, seq))

        add("Alpha synthase", core + "ACDEFGHIKLMNPQRS" * (gi + 1))          # lengths differ per genome
        add("Alpha synthase # frameshift", core[:20] + "WWWWWWWWWW")            # comment stripped, same function
        add("Beta kinase", other + core[:12])                                   # shares k-mers with Alpha: 80 % rule
        add("Beta kinase", other[:10] + "X" + other[11:] + "acdefghikl")        # X kills windows, lower case is valid
        add("Gamma lyase", "ACDEFGH")                                           # shorter than K: nothing
        add(None, core)                                                         # no assignment: skipped, no seq_id
        add("Rare thing %d" % gi, core[5:25])                                   # function in one genome only: not kept, still a seq_id
        add("Delta ligase", ("QWERTYIPASDFGHKLCVNM" * 4)[: 60 + gi])
        add("hypothetical protein", other + "TTTTTTTTTTTT")
        for rep in range(70):                                                    # 280 copies job-wide: the length sum wraps
            add("Epsilon pump", "HHHHHHHHKKKKKKKK" + "ACDEFGHIKLMNPQRSTVWY" * (10 + (rep * 7 + gi) % 9))
        genomes[g] = recs
    tree = str(tmp_path / "tree")
    write_tree(tree, genomes)
    deleted = str(tmp_path / "deleted.txt")
    with open(deleted, "w") as f:
        f.write("fig|10001.1.peg.3\n")
    for dfile in ("", deleted):
        refres = run_reference(ref, tree, str(tmp_path / ("ref_out%d" % bool(dfile))), deleted=dfile)
        table = run_oracle_through_dropin(tree, tmp_path / ("our_out%d" % bool(dfile)), str(tmp_path), deleted=dfile)
        assert_same(refres, table, "edge cases, deleted=%r" % dfile)
        kmers = refres[0]
        assert "HHHHHHHH" in kmers and "acdefghi" in kmers and not any("X" in k for k in kmers)


@pytest.mark.gpu
def test_gpu_command_line_equals_reference_sources(ref, tmp_path):
    """The whole drop-in (host gating + GPU build) against the reference's own builder on the same tree:
    every row and column of the kept table (median and var included: the reference runs single-threaded here,
    i.e. in canonical order), the printed counters and the per-function statistics."""
    s = Synth(n_proteins=4000, n_functions=120, n_genomes=6, seed=64)
    tree = str(tmp_path / "tree")
    s.write_tree(tree)
    kmers, cols, counters, df, swf = run_reference(ref, tree, str(tmp_path / "ref_out"))
    out = tmp_path / "gpu_out"
    r = subprocess.run([os.path.join(PKG, "kmers-build-signatures"), "-D", os.path.join(tree, "Annotations", "0"), "-F", os.path.join(tree, "Seqs"),
                        "--kmer-data-dir", str(out), "--final-kmers", "final.kmers", "--sorted-files", "--sigk-table", "kmer_data.sigk", "--no-recall"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    gk, gcols = read_table(out / "kmer_data.sigk")
    assert gk == kmers
    for name, a, b in zip(("avg_from_end", "function_index", "mean", "median", "var"), gcols, cols):
        np.testing.assert_array_equal(a, b, err_msg=name)
    assert "Kept %d kmers" % counters[0] in r.stdout
    assert "distinct_signatures=%d" % counters[1] in r.stdout
    assert "num_seqs_with_a_signature=%d" % counters[2] in r.stdout
    got_df = {int(l.split("\t")[0]): int(l.split("\t")[2]) for l in open(out / "distinct_functions").read().splitlines()}
    assert got_df == {i: int(c) for i, c in enumerate(df) if c}
    rows = [l.split("\t") for l in open(out / "final.kmers").read().splitlines()]
    assert [x[0] for x in rows] == kmers and [int(x[1]) for x in rows] == [int(v) for v in cols[0]]


# ---- the consumer side: the reference's FunctionCaller<KeptKmerDB<8>> from its own sources --------------
REF_CALL_SO = os.path.join(ROOT, "oracle", "_ref", "libref_call.so")


@pytest.fixture(scope="module")
def ref_call():
    if not os.path.exists(REF_CALL_SO):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], capture_output=True)
    if not os.path.exists(REF_CALL_SO):
        pytest.skip("oracle/_ref/libref_call.so not built (reference checkout absent)")
    lib = C.CDLL(REF_CALL_SO)
    u16p = np.ctypeslib.ndpointer(dtype=np.uint16, flags="C_CONTIGUOUS")
    lib.ref_call_functions.argtypes = [C.c_ulonglong, C.c_char_p, u16p, u16p, u16p, u16p, u16p, C.c_char_p, C.c_char_p, C.c_ulonglong,
                                       C.c_int, C.c_int, C.c_char_p, C.c_ulonglong]
    lib.ref_call_functions.restype = C.c_ulonglong
    return lib


def ref_calls(lib, table, names, fasta, tmp_path, ignore_hypo=False, want_calls=True):
    fi = os.path.join(str(tmp_path), "function.index")
    with open(fi, "w") as f:
        for i, name in enumerate(names):
            f.write("%d\t%s\n" % (i, name))
    data = fasta.encode("latin-1")
    out = C.create_string_buffer(1 << 20)
    n = lib.ref_call_functions(len(table.kmers), table.blob, *table.cols, fi.encode(), data, len(data), int(ignore_hypo), int(want_calls),
                               out, len(out))
    assert n <= len(out)
    return out.raw[:n].decode("latin-1").splitlines()


def test_function_caller_equals_reference_sources(ref_call, tmp_path):
    """host/function_caller.h against the reference's call_functions.tcc compiled from source: region calls and
    best calls on the hand-made cases of tests/test_function_caller.py and on tables built by the oracle."""
    import random

    import tests.test_function_caller as tfc

    host = tfc.load_host()
    ALPHA, BETA, NAMES = tfc.ALPHA, tfc.BETA, tfc.NAMES
    cases = []
    # single function, fragments, nothing known
    t = tfc.table_from({0: [ALPHA]}, {0: len(ALPHA)})
    cases.append((t, NAMES, [("p1", ALPHA), ("frag", ALPHA[:12]), ("tiny", ALPHA[:11]), ("none", "W" * 40), ("amb", ALPHA[:40] + "X" + ALPHA[41:])]))
    # chimeras: clear winner, and too close to call
    chim, chim2 = ALPHA[:60] + BETA[:40], ALPHA[:50] + BETA[:48]
    cases.append((tfc.table_from({0: [ALPHA], 1: [BETA]}, {0: len(chim), 1: len(chim)}), NAMES, [("c", chim)]))
    cases.append((tfc.table_from({0: [ALPHA], 1: [BETA]}, {0: len(chim2), 1: len(chim2)}), NAMES, [("c2", chim2)]))
    # the fusion-positive construction
    a, w, b = ALPHA[:80], ALPHA[100:160] + BETA[200:261], BETA[:90]
    rows = {}
    for p in range(len(a) - 7):
        rows[a[p:p + 8]] = (len(a) - p, 0, 280, 0, 0)
    for p in range(len(w) - 7):
        rows[w[p:p + 8]] = (len(w) - p, 2, 291 if p % 2 == 0 else 869, 0, 0)
    for p in range(len(b) - 7):
        rows[b[p:p + 8]] = (len(b) - p, 1, 300, 0, 0)
    cases.append((tfc.Table(rows), NAMES, [("fus", a + w + b)]))
    # random region lists
    rng = random.Random(19)
    for trial in range(40):
        q, rows = "", {}
        for r in range(rng.randrange(1, 7)):
            fi = rng.choice([0, 1, 2, 4])
            seq = "".join(rng.choice(tfc.AA) for _ in range(rng.randrange(12, 40)))
            for p in range(len(seq) - 7):
                rows[seq[p:p + 8]] = (len(seq) - p, fi, 0, 0, 0)
            q += seq
        rows = {k: (v[0], v[1], len(q), 0, 0) for k, v in rows.items()}
        cases.append((tfc.Table(rows), NAMES, [("t%d" % trial, q)]))
    # tables from the build oracle, training proteins and noisy chimeras
    for seed in (21, 22):
        t, seqs, funcs = tfc.build_table_with_oracle(seed, n_families=25, members=(3, 12), length=(60, 260), sub_rate=0.08, n_functions=12)
        names = ["Function %d" % i for i in range(12)]
        names[5] = "hypothetical protein"
        names[7] = "Function 3 / Function 4"
        recs = [("train%d" % i, (s.decode("latin-1") if isinstance(s, bytes) else s)) for i, s in enumerate(seqs[:60])]
        for i in range(30):
            x, y = rng.choice(recs)[1], rng.choice(recs)[1]
            q = x[:rng.randrange(10, len(x))] + y[rng.randrange(0, len(y) - 9):]
            recs.append(("chim%d" % i, "".join(c if rng.random() > 0.03 else rng.choice(tfc.AA + "X*") for c in q)))
        cases.append((t, names, recs))
    n_lines = 0
    for table, names, recs in cases:
        fasta = tfc.to_fasta(recs)
        for hypo in (False, True):
            want = ref_calls(ref_call, table, names, fasta, tmp_path, ignore_hypo=hypo)
            got = tfc.cxx_calls(host, table, names, fasta, ignore_hypo=hypo)
            assert got == want, recs[0][0]
            n_lines += len(want)
    assert n_lines > 500


def test_function_map_conventions_against_reference_sources(ref, tmp_path):
    """The input conventions of FunctionMap (src/function_map.h:62-332) on a deliberately messy tree: assignment
    comments and truncation markers, functions given on FASTA definition lines with [genome] brackets, ids that
    are not fig ids, explicit assignments overriding definition lines, multi-role functions with the good-roles
    list, the good-functions list, ignored functions, functions below the genome threshold."""
    aa = "ACDEFGHIKLMNPQRSTVWY"

    def seq(tag, n):
        return "".join(aa[(i * 7 + tag * 3 + (i // 5)) % 20] for i in range(n))

    root = str(tmp_path / "tree")
    os.makedirs(os.path.join(root, "Annotations", "0"))
    os.makedirs(os.path.join(root, "Seqs"))
    for gi in range(4):
        g = "2000%d.1" % gi
        ann, fa = [], []
        # plain assignments with comments; a truncated one is dropped
        ann.append(("fig|%s.peg.1" % g, "Alpha synthase (EC 1.1.1.1) # some note"))
        fa.append(("fig|%s.peg.1" % g, "", seq(1, 90 + gi)))
        ann.append(("fig|%s.peg.2" % g, "Alpha synthase (EC 1.1.1.1) ## another note"))
        fa.append(("fig|%s.peg.2" % g, "", seq(1, 80)))
        ann.append(("fig|%s.peg.3" % g, "Beta kinase # truncated"))
        fa.append(("fig|%s.peg.3" % g, "", seq(2, 70)))
        ann.append(("fig|%s.peg.4" % g, "Beta kinase # fragment"))
        fa.append(("fig|%s.peg.4" % g, "", seq(2, 75)))
        # multi-role functions: only genome 0 and 1 have them (below 3 genomes) -> kept through good roles
        if gi < 2:
            ann.append(("fig|%s.peg.5" % g, "Gamma lyase / Delta ligase"))
            fa.append(("fig|%s.peg.5" % g, "", seq(3, 100)))
            ann.append(("fig|%s.peg.6" % g, "Epsilon pump @ Zeta channel"))
            fa.append(("fig|%s.peg.6" % g, "", seq(4, 100)))
            ann.append(("fig|%s.peg.7" % g, "Eta factor; Theta factor"))
            fa.append(("fig|%s.peg.7" % g, "", seq(5, 100)))
            ann.append(("fig|%s.peg.8" % g, "Listed function"))             # kept through good functions
            fa.append(("fig|%s.peg.8" % g, "", seq(6, 100)))
            ann.append(("fig|%s.peg.9" % g, "Rare unlisted function"))      # dropped
            fa.append(("fig|%s.peg.9" % g, "", seq(7, 100)))
        # function on the definition line only, explicit assignment overriding a definition line
        fa.append(("fig|%s.peg.10" % g, " Iota reductase", seq(8, 85)))
        ann.append(("fig|%s.peg.11" % g, "Kappa oxidase"))
        fa.append(("fig|%s.peg.11" % g, " Something else entirely", seq(9, 85)))
        ann.append(("fig|%s.peg.12" % g, "Ignored function"))
        fa.append(("fig|%s.peg.12" % g, "", seq(10, 85)))
        fa.append(("fig|%s.peg.13" % g, "", seq(11, 85)))                   # no function anywhere
        with open(os.path.join(root, "Annotations", "0", g), "w") as f:
            for rid, fn in ann:
                f.write("%s\t%s\n" % (rid, fn))
        with open(os.path.join(root, "Seqs", g), "w") as f:
            for rid, d, s in fa:
                f.write(">%s%s\n%s\n" % (rid, d, s))
    # genbank-style files: no fig ids, function and genome on the definition line
    for gi in range(3):
        with open(os.path.join(root, "Seqs", "genbank%d" % gi), "w") as f:
            f.write(">prot%d_a Lambda transferase [Some organism %d]\n%s\n" % (gi, gi, seq(12, 95)))
            f.write(">prot%d_b Lambda transferase # truncated [Some organism %d]\n%s\n" % (gi, gi, seq(12, 60)))
            f.write(">prot%d_c Mu isomerase # a comment [Some organism %d]\n%s\n" % (gi, gi, seq(13, 95)))
    good_functions = str(tmp_path / "good_functions.txt")
    good_roles = str(tmp_path / "good_roles.txt")
    ignored = str(tmp_path / "ignored.txt")
    open(good_functions, "w").write("Listed function\n")
    open(good_roles, "w").write("Delta ligase\nZeta channel\nTheta factor\n")
    open(ignored, "w").write("Ignored function\n")

    out_ref = str(tmp_path / "ref_out")
    os.makedirs(out_ref)
    counters = (C.c_ulonglong * 3)()
    df = (C.c_uint * 65536)()
    swf = (C.c_uint * 65536)()
    ref.ref_signature_build_ex.argtypes = [C.c_char_p] * 6 + [C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_ulonglong), C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
    rc = ref.ref_signature_build_ex(os.path.join(root, "Annotations", "0").encode(), os.path.join(root, "Seqs").encode(), b"",
                                    good_functions.encode(), good_roles.encode(), ignored.encode(), 3, 1, out_ref.encode(), counters, df, swf)
    assert rc == 0
    kmers, cols = read_table(os.path.join(out_ref, "ref_table.bin"))
    refres = (kmers, cols, list(counters), np.array(df), np.array(swf))

    out = tmp_path / "our_out"
    dump = str(tmp_path / "packed.bin")
    r = subprocess.run([os.path.join(PKG, "kmers-build-signatures"), "-D", os.path.join(root, "Annotations", "0"), "-F", os.path.join(root, "Seqs"),
                        "--kmer-data-dir", str(out), "--good-functions", good_functions, "--good-roles", good_roles,
                        "--ignored-functions-file", ignored, "--sorted-files", "--dump-packed", dump], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    res, starts, func, sid = read_packed(dump)
    table, _ = oracle_c.oracle_build(PackedProteins(res.copy(), starts.copy(), func.copy(), sid.copy()))
    names_ref = function_index_names(os.path.join(out_ref, "function.index"))
    assert open(os.path.join(out_ref, "function.index")).read() == open(out / "function.index").read()
    kept = {n for _, n in names_ref}
    assert {"Alpha synthase (EC 1.1.1.1)", "Gamma lyase / Delta ligase", "Epsilon pump @ Zeta channel", "Eta factor; Theta factor",
            "Listed function", "Iota reductase", "Kappa oxidase", "hypothetical protein"} <= kept, kept
    assert "Rare unlisted function" not in kept and "Ignored function" not in kept and "Something else entirely" not in kept
    assert_same(refres, table, "messy tree")


def test_two_tier_contract_holds_for_the_reference_sources(ref, tmp_path):
    """SURVEY.md 8c's two tiers, checked on the reference's own code: feeding the same genomes in another file order
    leaves (k-mer, function_index, avg_from_end, mean) and every counter unchanged (tier A), while median / var —
    recurrences over the insertion order — may change (tier B is only defined for the canonical order)."""
    import shutil

    s = Synth(n_proteins=1200, n_functions=30, n_genomes=5, seed=65, zipf_s=1.1, mut_rate=0.03)
    tree_a = str(tmp_path / "a")
    s.write_tree(tree_a)
    # the same files under names that sort in reverse order (ids inside are untouched: the genome of a file comes
    # from its first fig id, src/function_map.h:167-175)
    tree_b = str(tmp_path / "b")
    for sub in (("Annotations", "0"), ("Seqs",)):
        src = os.path.join(tree_a, *sub)
        dst = os.path.join(tree_b, *sub)
        os.makedirs(dst)
        names = sorted(os.listdir(src))
        for i, name in enumerate(names):
            shutil.copy(os.path.join(src, name), os.path.join(dst, "%02d_%s" % (len(names) - i, name)))
    ka, ca, cnt_a, df_a, swf_a = run_reference(ref, tree_a, str(tmp_path / "out_a"))
    kb, cb, cnt_b, df_b, swf_b = run_reference(ref, tree_b, str(tmp_path / "out_b"))
    assert ka == kb and cnt_a == cnt_b
    for c in (0, 1, 2):                                         # avg_from_end, function_index, mean
        np.testing.assert_array_equal(ca[c], cb[c])
    np.testing.assert_array_equal(df_a, df_b)
    np.testing.assert_array_equal(swf_a, swf_b)
    differing = int((ca[3] != cb[3]).sum() + (ca[4] != cb[4]).sum())
    assert differing > 0, "median/var did not depend on the insertion order on this input"


def test_sequence_id_collisions_against_reference_sources(ref, tmp_path):
    """seq_id = file_number * max_seqs_per_file + n (src/signature_build.tcc:91,138): with more proteins in a file
    than max_seqs_per_file the ids of neighbouring files collide, and num_seqs_with_a_signature — the size of a
    set of ids (:274) — counts them once.  The constant is 100 000 in the command line; 7 here."""
    s = Synth(n_proteins=400, n_functions=10, n_genomes=4, seed=66)
    tree = str(tmp_path / "tree")
    s.write_tree(tree)
    ref.ref_set_max_seqs_per_file.argtypes = [C.c_int]
    ref.ref_set_max_seqs_per_file(7)
    try:
        refres = run_reference(ref, tree, str(tmp_path / "ref_out"))
    finally:
        ref.ref_set_max_seqs_per_file(100000)
    out = tmp_path / "our_out"
    dump = str(tmp_path / "packed.bin")
    r = subprocess.run([os.path.join(PKG, "kmers-build-signatures"), "-D", os.path.join(tree, "Annotations", "0"), "-F", os.path.join(tree, "Seqs"),
                        "--kmer-data-dir", str(out), "--sorted-files", "--max-seqs-per-file", "7", "--dump-packed", dump], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    res, starts, func, sid = read_packed(dump)
    assert len(set(sid.tolist())) < len(sid)                    # ids do collide
    table, _ = oracle_c.oracle_build(PackedProteins(res.copy(), starts.copy(), func.copy(), sid.copy()))
    assert_same(refres, table, "colliding sequence ids")
    assert table.num_seqs_with_a_signature == len(set(sid.tolist())) < 400


def test_float_threshold_divergence_against_reference_sources(ref, tmp_path):
    """The 80 % rule is `(float) best < float(count) * 0.8f` (src/signature_build.tcc:250-257).  It equals the exact
    5*best < 4*count below count = 10 485 764; there float(count) * 0.8f rounds DOWN to 8 388 611, so a k-mer with
    best = 8 388 611 of 10 485 764 occurrences is kept although 5*best < 4*count.  Two homopolymer proteins give
    exactly that group (and offsets that wrap 16 bits, a length sum that wraps, and a P-square walk of 8.4 M samples)."""
    tree = str(tmp_path / "tree")
    os.makedirs(os.path.join(tree, "Annotations", "0"))
    os.makedirs(os.path.join(tree, "Seqs"))
    best, count = 8_388_611, 10_485_764
    assert 5 * best < 4 * count
    with open(os.path.join(tree, "Annotations", "0", "5000.1"), "w") as f:
        f.write("fig|5000.1.peg.1\tMajor function\nfig|5000.1.peg.2\tMinor function\n")
    with open(os.path.join(tree, "Seqs", "5000.1"), "w") as f:
        f.write(">fig|5000.1.peg.1\n" + "A" * (best + 7) + "\n>fig|5000.1.peg.2\n" + "A" * (count - best + 7) + "\n")
    good = str(tmp_path / "good.txt")
    open(good, "w").write("Major function\nMinor function\n")
    out_ref = str(tmp_path / "ref_out")
    os.makedirs(out_ref)
    counters = (C.c_ulonglong * 3)()
    df = (C.c_uint * 65536)()
    swf = (C.c_uint * 65536)()
    ref.ref_signature_build_ex.argtypes = [C.c_char_p] * 6 + [C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_ulonglong), C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
    assert ref.ref_signature_build_ex(os.path.join(tree, "Annotations", "0").encode(), os.path.join(tree, "Seqs").encode(), b"", good.encode(),
                                      b"", b"", 3, 1, out_ref.encode(), counters, df, swf) == 0
    kmers, cols = read_table(os.path.join(out_ref, "ref_table.bin"))
    assert kmers == ["AAAAAAAA"]                                   # kept by the float test
    out = tmp_path / "our_out"
    dump = str(tmp_path / "packed.bin")
    r = subprocess.run([os.path.join(PKG, "kmers-build-signatures"), "-D", os.path.join(tree, "Annotations", "0"), "-F", os.path.join(tree, "Seqs"),
                        "--kmer-data-dir", str(out), "--good-functions", good, "--sorted-files", "--dump-packed", dump], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    res, starts, func, sid = read_packed(dump)
    table, _ = oracle_c.oracle_build(PackedProteins(res.copy(), starts.copy(), func.copy(), sid.copy()))
    assert_same((kmers, cols, list(counters), np.array(df), np.array(swf)), table, "float threshold at 10 485 764")
    # one occurrence fewer for the major function: rejected by both forms
    assert table.n_occurrences == count


@pytest.mark.gpu
def test_gpu_float_threshold_divergence(tmp_path):
    """The same group on the GPU: count = 10 485 764 >= 2^20 takes the float form of the keep rule, the group is
    reduced by the whole-warp walks and its median / var by the long order-statistics kernel.  (Last GPU test of the
    suite on purpose: the biggest single group any test builds.)"""
    from signature_kmers_b200.builder import GpuSignatureBuilder
    from tests.util import assert_tables_equal, pack

    best, count = 8_388_611, 10_485_764
    p = pack([b"A" * (best + 7), b"A" * (count - best + 7), b"A" * 20, b"CCCCCCCCCC"], [0, 1, 1, 2])
    b = GpuSignatureBuilder(device=0)
    b.set_proteins(p)
    got = b.build()
    want, _ = oracle_c.oracle_build(p)
    assert_tables_equal(got, want, tier_b=True, what="float threshold")
    # 13 more minor occurrences: 8 388 611 of 10 485 777 is rejected by both forms
    assert got.row("AAAAAAAA") is None and got.row("CCCCCCCC") is not None
    p2 = pack([b"A" * (best + 7), b"A" * (count - best + 7)], [0, 1])
    b.set_proteins(p2)
    got2 = b.build()
    want2, _ = oracle_c.oracle_build(p2)
    assert_tables_equal(got2, want2, tier_b=True, what="float threshold, exact divergence point")
    assert got2.row("AAAAAAAA")["function_index"] == 0
    b.close()


@pytest.mark.parametrize("seed", list(range(300, 340)))
def test_random_trees_against_reference_sources(ref, tmp_path, seed):
    """Small random protein sets (tests/util.random_proteins: shared domains, mixed functions around the 80 % boundary,
    ambiguity codes, lower case, ragged lengths, tiny alphabets for heavy duplication) written as trees: oracle ==
    reference sources on every column and counter."""
    from tests.util import random_proteins

    rng = np.random.default_rng(seed)
    alphabet = [b"ACDEFGHIKLMNPQRSTVWY", b"ACDEF", b"AC"][seed % 3]
    seqs, funcs = random_proteins(seed, n_families=int(rng.integers(3, 25)), members=(1, int(rng.integers(2, 30))),
                                  length=(5, int(rng.integers(20, 200))), sub_rate=float(rng.choice([0.0, 0.02, 0.1, 0.3])),
                                  n_functions=int(rng.integers(1, 9)), alphabet=alphabet)
    n_genomes = int(rng.integers(1, 5))
    genomes = {"4000%d.1" % g: [] for g in range(n_genomes)}
    for i, (s, f) in enumerate(zip(seqs, funcs)):
        g = "4000%d.1" % (i % n_genomes)
        fn = None if rng.random() < 0.05 else "Function number %d" % f
        genomes[g].append(("fig|%s.peg.%d" % (g, len(genomes[g]) + 1), fn, s.decode("latin-1")))
    tree = str(tmp_path / "tree")
    write_tree(tree, genomes)
    good = str(tmp_path / "good.txt")
    open(good, "w").write("".join("Function number %d\n" % f for f in range(0, 9, 2)))     # even ones kept whatever the evidence
    out_ref = str(tmp_path / "ref_out")
    os.makedirs(out_ref)
    counters = (C.c_ulonglong * 3)()
    df = (C.c_uint * 65536)()
    swf = (C.c_uint * 65536)()
    ref.ref_signature_build_ex.argtypes = [C.c_char_p] * 6 + [C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_ulonglong), C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
    min_reps = int(rng.integers(1, 4))
    assert ref.ref_signature_build_ex(os.path.join(tree, "Annotations", "0").encode(), os.path.join(tree, "Seqs").encode(), b"", good.encode(),
                                      b"", b"", min_reps, 1, out_ref.encode(), counters, df, swf) == 0
    kmers, cols = read_table(os.path.join(out_ref, "ref_table.bin"))
    out = tmp_path / "our_out"
    dump = str(tmp_path / "packed.bin")
    r = subprocess.run([os.path.join(PKG, "kmers-build-signatures"), "-D", os.path.join(tree, "Annotations", "0"), "-F", os.path.join(tree, "Seqs"),
                        "--kmer-data-dir", str(out), "--good-functions", good, "--min-reps-required", str(min_reps), "--sorted-files",
                        "--dump-packed", dump], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    res, starts, func, sid = read_packed(dump)
    table, _ = oracle_c.oracle_build(PackedProteins(res.copy(), starts.copy(), func.copy(), sid.copy()))
    assert_same((kmers, cols, list(counters), np.array(df), np.array(swf)), table, "random tree %d" % seed)
    assert open(os.path.join(out_ref, "function.index")).read() == open(out / "function.index").read()


# ---- the reference's whole command line (its main) over the stand-ins -------------------------------------
REF_MAIN = os.path.join(ROOT, "oracle", "_ref", "ref-kmers-build-signatures")


def host_outputs(tree, out, table, extra=None, min_reps=3, n_threads=2):
    """The drop-in's host code writes every output file from a table given from outside (libsigk_host.so)."""
    extra = extra or {}
    lib = C.CDLL(os.path.join(PKG, "libsigk_host.so"))
    u16p = np.ctypeslib.ndpointer(dtype=np.uint16, flags="C_CONTIGUOUS")
    u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
    lib.sigk_host_outputs.argtypes = [C.c_char_p] * 6 + [C.c_int, C.c_int, C.c_char_p, C.c_uint64, C.c_char_p, u16p, u16p, u16p, u16p, u16p, u32p]
    rc = lib.sigk_host_outputs(os.path.join(tree, "Annotations", "0").encode(), os.path.join(tree, "Seqs").encode(),
                               extra.get("good_functions", "").encode(), extra.get("good_roles", "").encode(),
                               extra.get("ignored", "").encode(), extra.get("deleted", "").encode(), min_reps, n_threads, str(out).encode(),
                               table.n_kept, np.ascontiguousarray(table.kmer).tobytes(), np.ascontiguousarray(table.avg_from_end),
                               np.ascontiguousarray(table.function_index), np.ascontiguousarray(table.mean), np.ascontiguousarray(table.median),
                               np.ascontiguousarray(table.var), np.ascontiguousarray(table.distinct_functions, dtype=np.uint32))
    assert rc == 0


@pytest.mark.parametrize("case", ["synthetic", "noisy", "messy"])
def test_every_output_file_against_the_reference_command_line(ref, tmp_path, case):
    """The reference's own main() (compiled over the stand-ins) and the drop-in's host code on the same tree: stdout
    banner and counters, function.index, otu.index, genomes, final.kmers (the reference writes hash order: compared as
    sorted lines), distinct_functions (likewise), and recall.report.d file by file, byte for byte.  The drop-in's table
    comes from the CPU oracle here; the GPU tests hold the GPU table to the same oracle."""
    if not os.path.exists(REF_MAIN):
        pytest.skip("oracle/_ref/ref-kmers-build-signatures not built (reference checkout absent)")
    tree = str(tmp_path / "tree")
    extra = {}
    if case == "synthetic":
        Synth(n_proteins=900, n_functions=30, n_genomes=5, seed=67).write_tree(tree)
    elif case == "noisy":          # some families never reach 3 genomes: their proteins are called differently or not at all
        Synth(n_proteins=330, n_functions=66, n_genomes=4, seed=68, zipf_s=0.0, mut_rate=0.3).write_tree(tree)
    else:
        import shutil
        shutil.copytree(os.path.join(ROOT, "tests", "golden", "edge", "tree"), tree)
        for key, name in (("good_functions", "good_functions.txt"), ("good_roles", "good_roles.txt"), ("ignored", "ignored.txt"), ("deleted", "deleted.txt")):
            extra[key] = str(tmp_path / name)
            os.replace(os.path.join(tree, name), extra[key])
    for junk in ("function.index.expected",):
        if os.path.exists(os.path.join(tree, junk)):
            os.remove(os.path.join(tree, junk))
    ref_out = tmp_path / "ref_out"
    cmd = [REF_MAIN, "-D", os.path.join(tree, "Annotations", "0"), "-F", os.path.join(tree, "Seqs"), "--kmer-data-dir", str(ref_out),
           "--final-kmers", "final.kmers", "--min-reps-required", "3", "--n-threads", "1"]
    ours = [os.path.join(PKG, "kmers-build-signatures")] + cmd[1:5] + ["--kmer-data-dir", str(tmp_path / "dump_out"), "--min-reps-required", "3"]
    for flag, key in (("--good-functions", "good_functions"), ("--good-roles", "good_roles"), ("--ignored-functions-file", "ignored"),
                      ("--deleted-features-file", "deleted")):
        if key in extra:
            cmd += [flag, extra[key]]
            ours += [flag, extra[key]]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    # the oracle's table for the same tree (packed proteins from the drop-in's host code, readdir order like the reference)
    dump = str(tmp_path / "packed.bin")
    d = subprocess.run(ours + ["--dump-packed", dump], capture_output=True, text=True)
    assert d.returncode == 0, d.stderr
    res, starts, func, sid = read_packed(dump)
    table, _ = oracle_c.oracle_build(PackedProteins(res.copy(), starts.copy(), func.copy(), sid.copy()))
    our_out = tmp_path / "our_out"
    host_outputs(tree, our_out, table, extra)

    ref_lines = r.stdout.splitlines()
    assert d.stdout.splitlines()[:4] == ref_lines[:4]                       # definitions / fasta / keep / kept N functions
    assert ref_lines[4:7] == ["Kept %d kmers" % table.n_kept, "distinct_signatures=%d" % table.distinct_signatures,
                              "num_seqs_with_a_signature=%d" % table.num_seqs_with_a_signature]
    for name in ("function.index", "otu.index", "genomes"):
        assert open(ref_out / name).read() == open(our_out / name).read(), name
    # (the reference writes hash order, the drop-in table order: compared as sorted lines)
    assert sorted(open(ref_out / "final.kmers").read().splitlines()) == sorted(open(our_out / "final.kmers").read().splitlines())
    assert sorted(open(ref_out / "distinct_functions").read().splitlines(), key=lambda l: int(l.split("\t")[0])) == \
        open(our_out / "distinct_functions").read().splitlines()
    ref_reports = {f: open(ref_out / "recall.report.d" / f).read() for f in os.listdir(ref_out / "recall.report.d")}
    our_reports = {f: open(our_out / "recall.report.d" / f).read() for f in os.listdir(our_out / "recall.report.d")}
    assert ref_reports == our_reports
    if case == "messy":
        assert sum(len(v) for v in ref_reports.values()) > 0                # un-kept, ignored and truncated assignments are reported


def test_config1_kept_table_equals_reference_sources(ref, tmp_path):
    """BASELINE.json configs[0] at its full size (20 K proteins / 1 K functions, 5.9 M occurrences, 4.4 M kept
    k-mers): the kept table of the reference's own sources, every row and column, against the CPU oracle fed by the
    drop-in's host code.  (About 20 s: the stand-in containers are serial.)"""
    tree = str(tmp_path / "tree")
    Synth.config("config1").write_tree(tree)
    refres = run_reference(ref, tree, str(tmp_path / "ref_out"))
    table = run_oracle_through_dropin(tree, tmp_path / "our_out", str(tmp_path))
    assert table.n_kept > 4_000_000
    assert_same(refres, table, "config1")
    assert open(tmp_path / "ref_out" / "function.index").read() == open(tmp_path / "our_out" / "function.index").read()


def test_function_caller_fuzz_against_reference_sources(ref_call, tmp_path):
    """Random hit patterns through the whole call state machine: regions of 1..40 hits, spacers longer than max_gap
    (200), functions alternating hit by hit, k-mers whose stored lengths vary (so the median absolute deviation is not
    the default 30) or sit far from the query length (regions that fail their length test), ambiguity codes inside
    regions, the fusion naming.  host/function_caller.h == the reference's call_functions.tcc, line by line."""
    import random

    import tests.test_function_caller as tfc

    host = tfc.load_host()
    names = ["Alpha synthase", "Beta kinase", "Alpha synthase / Beta kinase", "hypothetical protein", "Gamma lyase", "Delta ligase / Gamma lyase"]
    rng = random.Random(77)
    n_lines = n_calls = 0
    for trial in range(800):
        rows, q = {}, ""
        qlen_guess = rng.choice([120, 300, 600])
        for seg in range(rng.randrange(1, 10)):
            kind = rng.random()
            if kind < 0.25:                                           # spacer: residues that are in no table k-mer
                q += "".join(rng.choice("wy") for _ in range(rng.choice([0, 3, 9, 150, 199, 200, 201, 260])))
                continue
            fi = rng.randrange(len(names))
            length = rng.randrange(8, 48)
            seq = "".join(rng.choice(tfc.AA) for _ in range(length))
            if kind > 0.9 and length > 20:
                seq = seq[:10] + rng.choice("X*") + seq[11:]
            spread = rng.choice([0, 0, 5, 40, 400])
            for p in range(len(seq) - 7):
                k = seq[p:p + 8]
                if "X" in k or "*" in k:
                    continue
                f2 = fi if rng.random() > 0.15 else rng.randrange(len(names))     # a stray hit for another function
                rows[k] = (len(seq) - p, f2, max(0, min(65535, qlen_guess + rng.randint(-spread, spread))), 0, 0)
            q += seq
        if not rows:
            continue
        table = tfc.Table(rows)
        fasta = tfc.to_fasta([("q%d" % trial, q)])
        for hypo in (False, True):
            want = ref_calls(ref_call, table, names, fasta, tmp_path, ignore_hypo=hypo)
            got = tfc.cxx_calls(host, table, names, fasta, ignore_hypo=hypo)
            assert got == want, (trial, q)
            assert tfc.oracle_calls(table, names, [("q%d" % trial, q)], ignore_hypo=hypo) == want, (trial, q)    # the Python restatement too
            n_lines += len(want)
            n_calls += sum(1 for l in want if l.startswith("#call"))
    assert n_lines > 500 and n_calls > 100


@pytest.mark.parametrize("seed", list(range(500, 620)))
def test_assignment_text_fuzz_against_reference_sources(ref, tmp_path, seed):
    """Random assignment and definition-line texts built from the pieces the SEED conventions care about (comment
    markers with and without blanks, truncation words, role separators, [genome] brackets, tabs, repeated blanks):
    kept functions, their indices, function.index and the kept table of the reference's sources vs the drop-in's
    hand-matched text handling (host/seed_text.h, FunctionMap)."""
    import random

    rng = random.Random(seed)
    bases = ["Alpha synthase", "Beta kinase (EC 2.7.1.1)", "Gamma lyase", "Delta ligase", "hypothetical protein", "Epsilon pump subunit B",
             "Zeta-channel protein", "Eta factor"]
    seps = [" / ", " @ ", "; ", " /", "/ ", " ; ", " @", "@"]
    comments = ["", "", "", " # truncated", " # fragment", " # frameshift", " ## missing start", " #truncated", "# note", "  #  trunc at end",
                " ! comment", " # ", " #", " ### x", " # Truncated", " # frag"]

    def function_text():
        f = rng.choice(bases)
        if rng.random() < 0.3:
            f += rng.choice(seps) + rng.choice(bases)
        if rng.random() < 0.1:
            f = " " + f
        if rng.random() < 0.1:
            f += " "
        return f + rng.choice(comments)

    aa = "ACDEFGHIKLMNPQRSTVWY"
    protein = {b: "".join(rng.choice(aa) for _ in range(rng.randrange(30, 90))) for b in bases}
    tree = str(tmp_path / "tree")
    os.makedirs(os.path.join(tree, "Annotations", "0"))
    os.makedirs(os.path.join(tree, "Seqs"))
    for gi in range(rng.randrange(2, 5)):
        g = "6000%d.1" % gi
        with open(os.path.join(tree, "Annotations", "0", g), "w") as ann, open(os.path.join(tree, "Seqs", g), "w") as fa:
            for n in range(1, rng.randrange(8, 25)):
                rid = "fig|%s.peg.%d" % (g, n) if rng.random() > 0.1 else "other%d_%d" % (gi, n)
                text = function_text()
                base = next(b for b in bases if b in text)
                seq = "".join(c if rng.random() > 0.05 else rng.choice(aa) for c in protein[base])
                how = rng.random()
                definition = ""
                if how < 0.6:
                    ann.write("%s\t%s\n" % (rid, text))
                elif how < 0.8:
                    definition = " " + function_text() + (" [Genome %d.%d]" % (gi, n % 3) if rng.random() < 0.6 else "")
                elif how < 0.9:
                    ann.write("%s\t%s\n" % (rid, text))
                    definition = rng.choice([" ", "  ", "\t"]) + function_text() + rng.choice(["", " [g]", " [a] [b]", " []", " [x]y"])
                if rng.random() < 0.05:
                    ann.write("%s\t%s\textra column\n" % (rid, function_text()))          # a second assignment for the same id
                fa.write(">%s%s\n%s\n" % (rid, definition, seq))
    good_roles = str(tmp_path / "good_roles.txt")
    open(good_roles, "w").write("Gamma lyase\nEta factor\n")
    out_ref = str(tmp_path / "ref_out")
    os.makedirs(out_ref)
    counters = (C.c_ulonglong * 3)()
    df = (C.c_uint * 65536)()
    swf = (C.c_uint * 65536)()
    ref.ref_signature_build_ex.argtypes = [C.c_char_p] * 6 + [C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_ulonglong), C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
    assert ref.ref_signature_build_ex(os.path.join(tree, "Annotations", "0").encode(), os.path.join(tree, "Seqs").encode(), b"", b"",
                                      good_roles.encode(), b"", 2, 1, out_ref.encode(), counters, df, swf) == 0
    kmers, cols = read_table(os.path.join(out_ref, "ref_table.bin"))
    out = tmp_path / "our_out"
    dump = str(tmp_path / "packed.bin")
    r = subprocess.run([os.path.join(PKG, "kmers-build-signatures"), "-D", os.path.join(tree, "Annotations", "0"), "-F", os.path.join(tree, "Seqs"),
                        "--kmer-data-dir", str(out), "--good-roles", good_roles, "--min-reps-required", "2", "--sorted-files", "--dump-packed", dump],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(os.path.join(out_ref, "function.index")).read() == open(out / "function.index").read()
    res, starts, func, sid = read_packed(dump)
    table, _ = oracle_c.oracle_build(PackedProteins(res.copy(), starts.copy(), func.copy(), sid.copy()))
    assert_same((kmers, cols, list(counters), np.array(df), np.array(swf)), table, "assignment text fuzz %d" % seed)
