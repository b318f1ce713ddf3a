"""Function calling from the kept table: the C++ host code
(signature_kmers_b200/host/function_caller.h, through libsigk_host.so and the kmers-call-functions command
line) against the independent Python restatement oracle/call_oracle.py and hand-derived known answers
(expected values derived from /root/reference/src/call_functions.tcc and src/kmer_data.h:76-102)."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import call_oracle as co
from tests.util import pack, random_proteins
from signature_kmers_b200.capi import table_order_key

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "signature_kmers_b200")
AA = "ACDEFGHIKLMNPQRSTVWY"


def load_host():
    if not os.path.exists(os.path.join(PKG, "libsigk.so")):
        import __graft_entry__ as g
        g.build()
    subprocess.run(["make", "-C", os.path.join(PKG, "host"), "../libsigk_host.so", "../kmers-call-functions"], check=True, capture_output=True)
    lib = C.CDLL(os.path.join(PKG, "libsigk_host.so"))
    u16p = np.ctypeslib.ndpointer(dtype=np.uint16, flags="C_CONTIGUOUS")
    lib.sigk_host_call_functions.argtypes = [C.c_uint64, C.c_char_p, u16p, u16p, u16p, u16p, u16p, C.c_char_p, C.c_char_p, C.c_uint64,
                                             C.c_int, C.c_int, C.c_char_p, C.c_uint64]
    lib.sigk_host_call_functions.restype = C.c_uint64
    lib.sigk_host_call_windows.argtypes = [C.c_char_p, C.c_uint64, np.ctypeslib.ndpointer(dtype=np.uint32), C.c_uint64]
    lib.sigk_host_call_windows.restype = C.c_uint64
    return lib


@pytest.fixture(scope="module")
def host():
    return load_host()


class Table:
    """A kept table given as {kmer: (avg_from_end, function_index, mean, median, var)}, in the layout libsigk returns."""

    def __init__(self, rows: dict):
        self.rows = rows
        self.kmers = sorted(rows, key=table_order_key)          # the table order of include/sigk.h
        self.blob = "".join(self.kmers).encode("latin-1")
        cols = np.array([rows[k] for k in self.kmers], dtype=np.uint16).reshape(-1, 5)
        self.cols = [np.ascontiguousarray(cols[:, i]) for i in range(5)]

    def write_sigk(self, path):
        with open(path, "wb") as f:
            f.write(b"SIGKTBL1")
            f.write(np.uint64(len(self.kmers)).tobytes())
            f.write(self.blob)
            for c in self.cols:
                f.write(c.tobytes())


def cxx_calls(host, table: Table, names, fasta: str, ignore_hypo=False, want_calls=True):
    data = fasta.encode("latin-1")
    out = C.create_string_buffer(1 << 16)
    args = [len(table.kmers), table.blob, *table.cols, "\n".join(names).encode(), data, len(data), int(ignore_hypo), int(want_calls)]
    n = host.sigk_host_call_functions(*args, out, len(out))
    if n > len(out):
        out = C.create_string_buffer(n)
        n = host.sigk_host_call_functions(*args, out, len(out))
    return out.raw[:n].decode("latin-1").splitlines()


def oracle_calls(table: Table, names, records, ignore_hypo=False, want_calls=True):
    fc = co.FunctionCaller(table.rows, names, ignore_hypothetical=ignore_hypo)
    lines = []
    for rid, seq in records:
        calls = fc.process_aa_seq(seq)
        if want_calls:
            for c in calls:
                lines.append("#call\t%d\t%d\t%d\t%d\t%d\t%s" % (c["start"], c["end"], c["count"], c["function_index"], c["median"],
                                                               co.format_score(c["mad"])))
        fi, func, score = fc.find_best_call(calls)
        lines.append("%s\t%s\t%d\t%s" % (rid, func, fi, co.format_score(score)))
    return lines


def to_fasta(records):
    return "".join(">%s some definition\n%s\n" % (rid, seq) for rid, seq in records)


# ---- for_each_kmer (src/kmer_data.h:76-102) ---------------------------------------------------------
def test_windows_known_answers(host):
    def cxx(seq):
        buf = np.zeros(len(seq) + 1, dtype=np.uint32)
        n = host.sigk_host_call_windows(seq.encode(), len(seq), buf, len(buf))
        return [int(x) for x in buf[:n]]

    # no ambiguity: every position 0 .. len-8
    assert cxx("ACDEFGHIKLMN") == [0, 1, 2, 3, 4]
    # X at 9: the window at 0 ends at 8 (< 9) and is visited; the window at 1 ENDS right before the X (kend = 9 >= 9)
    # and is skipped although it does not contain it; the scan resumes after the X
    assert cxx("ACDEFGHIKXLMNPQRSTVW") == [0, 10, 11, 12]
    # '*' behaves like X; lower case, B and Z are ordinary characters on the call side
    assert cxx("ACDEFGHI*ACDEFGHI") == [9]          # the window at 0 ends at 8 = the '*' position: skipped as well
    assert cxx("acdebzhiklm") == [0, 1, 2, 3]
    assert cxx("ACDEFGH") == [] and cxx("") == []
    # two ambiguities closer than a window: nothing between them
    assert cxx("ACDEFGHIKLXMNPQXRSTVWYACD") == [0, 1, 16, 17]
    for seq in ("ACDEFGHIKXLMNPQRSTVW", "XXXXXXXXXXXX", "ACDEFGHIX", "XACDEFGHI", "ACDEFGHIKLMNPQRSTVWY*"):
        assert cxx(seq) == [off for _, off in co.for_each_kmer(seq)]


def test_windows_random_against_oracle(host):
    rng = random.Random(5)
    for _ in range(300):
        n = rng.randrange(0, 60)
        seq = "".join(rng.choice(AA + "X*xb") for _ in range(n))
        buf = np.zeros(n + 1, dtype=np.uint32)
        k = host.sigk_host_call_windows(seq.encode(), n, buf, len(buf))
        assert [int(x) for x in buf[:k]] == [off for _, off in co.for_each_kmer(seq)], seq


# ---- known answers for the call logic ---------------------------------------------------------------
NAMES = ["Alpha synthase", "Beta kinase", "Alpha synthase / Beta kinase", "hypothetical protein", "Gamma lyase"]


def table_from(seq_by_func: dict, mean_by_func: dict) -> Table:
    rows = {}
    for fi, seqs in seq_by_func.items():
        for seq in seqs:
            for p in range(len(seq) - 7):
                rows[seq[p:p + 8]] = (len(seq) - p, fi, mean_by_func[fi], 0, 0)
    return Table(rows)


ALPHA = "MKTAYIAKQRQISFVKSHFSRQLEERLGLIEVQAPILSRVGDGTQDNLSGAEKAVQVKVKALPDAQFEVVHSLAKWKRQTLGQHDFSAGEGLYTHMKALRPDEDRLSPLHSVYVDQWDWERVMGDGERQFSTLKSTVEAIWAGIKATEAAVSEEFGLAPFLPDQIHFVHSQELLSRYPDLDAKGRERAIAKDLGAVFLVGIGGKLSDGHRHDVRAPDYDDWSTPSELGHAGLNGDILVWNPVLEDAFELSSMGIRVDADTLKHQLALTGDEDRLELEWHQALLRGEMPQTIGGGIGQSRLTMLLLQLPHIGQVQAGVWPAAVRESVPSLL"
BETA = "MSDNGELEDKPPAPPVRMSSTIFSTGGKDPLSANHSLKPLPSVPEEKKPRHKIISIFSGTEKGSKKKEKERPEISPPSDFEHTIHVGFDAVTGEFTGMPEQWARLLQTSNITKLEQKKNPQAVLDVLKFYDSNTVKQKYLSFTPPEKDGFPSGTPALNAKGTEAPAVVTEEEDDDEETAPPVIAPRPDHTKSIYTRSVIDPVPAPVGDSHVDGAAKSLDKQKKKTKMTDEEIMEKLRTIVSIGDPKKKYTRYEKIGQGASGTVFTATDVALGQEVAIKQINLQKQPKKELIINEILVMKELKNPNIVNFLDSYLVGDELFVVMEYLAGGSLTDVVTETCMDEAQIAAVCRECLQALEFLHANQVIHRDIKSDNVLLGMEGSVKLTDFGFCAQITPEQSKRSTMVGTPYWMAPEVVTRKAYGPKVDIWSLGIMAIEMVEGEPPYLNENPLRALYLIATNGTPELQNPEKLSPIFRDFLNRCLEMDVEKRGSAKELLQHPFLKLAKPLSSLTPLIMAAKEAMKSNR"


def test_single_function_call_known_answer(host):
    t = table_from({0: [ALPHA]}, {0: len(ALPHA)})
    # the protein itself: every window hits function 0; one region, count = len-7, length test passes (mad 0 -> 30)
    lines = cxx_calls(host, t, NAMES, to_fasta([("p1", ALPHA)]))
    n = len(ALPHA) - 7
    assert lines == ["#call\t0\t%d\t%d\t0\t%d\t30" % (len(ALPHA) - 1, n, len(ALPHA)), "p1\tAlpha synthase\t0\t%d" % n]
    # a fragment of 12 residues: 5 hits = min_hits, but its length is far below mean - 2*30 -> no region, no call
    assert cxx_calls(host, t, NAMES, to_fasta([("frag", ALPHA[:12])])) == ["frag\t\t65535\t0"]
    # 4 hits only (11 residues): below min_hits
    assert cxx_calls(host, t, NAMES, to_fasta([("tiny", ALPHA[:11])])) == ["tiny\t\t65535\t0"]
    # nothing known
    assert cxx_calls(host, t, NAMES, to_fasta([("none", "W" * 40)])) == ["none\t\t65535\t0"]


def test_two_functions_known_answers(host):
    la, lb = len(ALPHA), len(BETA)
    # a chimera of a 60-residue piece of ALPHA and a 40-residue piece of BETA, table means set to the chimera's length
    chim = ALPHA[:60] + BETA[:40]
    t = table_from({0: [ALPHA], 1: [BETA]}, {0: len(chim), 1: len(chim)})
    lines = cxx_calls(host, t, NAMES, to_fasta([("c", chim)]))
    # hits: 53 for function 0 (offsets 0..52), then 33 for function 1 (offsets 60..92).  The second function-1 hit
    # closes the first region (count 53) and the run restarts with those two hits.
    assert lines[0] == "#call\t0\t59\t53\t0\t%d\t30" % len(chim)
    assert lines[1] == "#call\t60\t99\t33\t1\t%d\t30" % len(chim)
    assert lines[2] == "c\tAlpha synthase\t0\t53"           # 53 - 33 >= 5
    # nearly balanced pieces: offset below 5 -> "larger name ?? smaller name", score of the best
    chim2 = ALPHA[:50] + BETA[:48]
    t2 = table_from({0: [ALPHA], 1: [BETA]}, {0: len(chim2), 1: len(chim2)})
    assert cxx_calls(host, t2, NAMES, to_fasta([("c2", chim2)]), want_calls=False) == ["c2\tBeta kinase ?? Alpha synthase\t65535\t43"]
    assert la and lb


def test_fusion_known_answer(host):
    # A | W | B with part lengths that add up: mean(A) + mean(B) ~ mean(W) -> the fusion function, score = all hits
    a, b = ALPHA[:80], BETA[:90]
    w = ALPHA[100:160] + BETA[200:260]
    q = a + w + b
    rows = {}
    for fi, seq, mean in ((0, a, 100), (2, w, 210), (1, b, 110)):
        for p in range(len(seq) - 7):
            rows[seq[p:p + 8]] = (len(seq) - p, fi, len(q), 0, 0)
    t = Table(rows)
    fc = co.FunctionCaller(t.rows, NAMES)
    calls = fc.process_aa_seq(q)
    assert [c["function_index"] for c in calls] == [0, 2, 1]
    # protein_length_median of every part = len(q): mean(A)+mean(B)-mean(W) = len(q) -> frac = 1: NOT a fusion
    assert cxx_calls(host, t, NAMES, to_fasta([("q", q)]), want_calls=False) == oracle_calls(t, NAMES, [("q", q)], want_calls=False)
    # now give the parts lengths that add up; the length test of each region needs |len(q) - mean| <= 2*30
    rows2 = {}
    for fi, seq, mean in ((0, a, len(q) - 40), (2, w, 2 * len(q) - 60), (1, b, len(q) - 20)):
        for p in range(len(seq) - 7):
            rows2[seq[p:p + 8]] = (len(seq) - p, fi, mean, 0, 0)
    t2 = Table(rows2)
    got = cxx_calls(host, t2, NAMES, to_fasta([("q", q)]), want_calls=False)
    assert got == oracle_calls(t2, NAMES, [("q", q)], want_calls=False)
    # the W region fails its own length test (2 len - 60 is far from len), so only A and B remain: no fusion call
    assert got == ["q\tBeta kinase\t1\t83"]


def test_fusion_positive_known_answer(host):
    # |A: 73 hits|W: 114 hits|B: 83 hits|, query length 291.  A's k-mers carry mean 280, B's 300 (both within
    # 291 +- 60 with mad 0 -> 30).  W's k-mers alternate between 291 and 869: mean 580, median (291+869)/2 = 580,
    # mad 289, so W's length window [2, 1158] holds 291.  The letters are "AWB" (fusion function = 'W'), and
    # |280 + 300 - 580| / 580 = 0 < 0.1: the compound function is called with the sum of all hits.
    a, w, b = ALPHA[:80], ALPHA[100:160] + BETA[200:261], BETA[:90]
    q = a + w + b
    assert len(q) == 291 and len(w) - 7 == 114
    rows = {}
    for p in range(len(a) - 7):
        rows[a[p:p + 8]] = (len(a) - p, 0, 280, 0, 0)
    for p in range(len(w) - 7):
        rows[w[p:p + 8]] = (len(w) - p, 2, 291 if p % 2 == 0 else 869, 0, 0)
    for p in range(len(b) - 7):
        rows[b[p:p + 8]] = (len(b) - p, 1, 300, 0, 0)
    t = Table(rows)
    got = cxx_calls(host, t, NAMES, to_fasta([("fus", q)]))
    assert got == ["#call\t0\t79\t73\t0\t280\t30", "#call\t80\t200\t114\t2\t580\t289", "#call\t201\t290\t83\t1\t300\t30",
                   "fus\tAlpha synthase / Beta kinase\t2\t270"]
    assert got == oracle_calls(t, NAMES, [("fus", q)])


def test_best_call_logic_against_oracle_on_crafted_calls(host):
    """find_best_call on hand-made region lists, driven through tables whose k-mers are unique per region."""
    rng = random.Random(9)
    names = NAMES
    for trial in range(60):
        n_regions = rng.randrange(1, 7)
        q = ""
        rows = {}
        for r in range(n_regions):
            fi = rng.choice([0, 1, 2, 4])
            length = rng.randrange(12, 40)
            seq = "".join(rng.choice(AA) for _ in range(length))
            for p in range(length - 7):
                rows[seq[p:p + 8]] = (length - p, fi, 0, 0, 0)
            q += seq
        # one mean for all rows so that every region passes the length test
        rows = {k: (v[0], v[1], len(q), 0, 0) for k, v in rows.items()}
        t = Table(rows)
        rec = [("t%d" % trial, q)]
        assert cxx_calls(host, t, names, to_fasta(rec)) == oracle_calls(t, names, rec), q


# ---- realistic tables: the CPU oracle of the build, then calls ---------------------------------------
def build_table_with_oracle(seed, **kw):
    from oracle import oracle_c

    seqs, funcs = random_proteins(seed, **kw)
    table, _ = oracle_c.oracle_build(pack(seqs, funcs))
    rows = {}
    for i, k in enumerate(table.kmer_strings()):
        rows[k] = (int(table.avg_from_end[i]), int(table.function_index[i]), int(table.mean[i]), int(table.median[i]), int(table.var[i]))
    return Table(rows), seqs, funcs


@pytest.mark.parametrize("seed", [21, 22, 23])
def test_calls_match_oracle_on_built_tables(host, seed):
    t, seqs, funcs = build_table_with_oracle(seed, n_families=25, members=(3, 12), length=(60, 260), sub_rate=0.08, n_functions=12)
    names = ["Function %d" % i for i in range(12)]
    names[5] = "hypothetical protein"
    names[7] = "Function 3 / Function 4"
    rng = random.Random(seed)
    records = []
    for i, s in enumerate(seqs[:80]):
        s = s.decode("latin-1") if isinstance(s, bytes) else s
        records.append(("train%d" % i, s))
    for i in range(40):             # chimeras, mutated copies, ambiguity codes, fragments
        a = rng.choice(seqs); b = rng.choice(seqs)
        a = a.decode("latin-1") if isinstance(a, bytes) else a
        b = b.decode("latin-1") if isinstance(b, bytes) else b
        q = a[:rng.randrange(10, len(a))] + b[rng.randrange(0, len(b) - 9):]
        q = "".join(c if rng.random() > 0.03 else rng.choice(AA + "X*") for c in q)
        records.append(("chim%d" % i, q))
    for hypo in (False, True):
        got = cxx_calls(host, t, names, to_fasta(records), ignore_hypo=hypo)
        want = oracle_calls(t, names, records, ignore_hypo=hypo)
        assert got == want
    called = [l for l in cxx_calls(host, t, names, to_fasta(records[:80]), want_calls=False) if not l.endswith("\t65535\t0")]
    assert len(called) > 40          # the training proteins are recalled


# ---- the command line ---------------------------------------------------------------------------------
def test_command_line(host, tmp_path):
    t, seqs, funcs = build_table_with_oracle(31, n_families=10, members=(4, 8), length=(80, 200), sub_rate=0.05, n_functions=6)
    names = ["F%d" % i for i in range(6)] + ["hypothetical protein"]
    data = tmp_path / "data"
    data.mkdir()
    t.write_sigk(data / "kmer_data.sigk")
    with open(data / "function.index", "w") as f:
        for i in reversed(range(len(names))):               # ids in any order, extra columns ignored
            f.write("%d\t%s\t3\t1.0\t1.0\t0\t0\n" % (i, names[i]))
    recs = [("fig|1.1.peg.%d" % i, (s.decode("latin-1") if isinstance(s, bytes) else s)) for i, s in enumerate(seqs[:30])]
    fa1, fa2 = tmp_path / "a.fa", tmp_path / "b.fa"
    fa1.write_text(to_fasta(recs[:20]))
    fa2.write_text(to_fasta(recs[20:]))
    exe = os.path.join(PKG, "kmers-call-functions")
    want = oracle_calls(t, names, recs, want_calls=False)
    r = subprocess.run([exe, str(data), str(fa1), str(fa2)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout.splitlines() == want and "Data size 10" in r.stderr
    out = tmp_path / "calls.txt"
    r = subprocess.run([exe, "-d", str(data), "-i", str(fa1), str(fa2), "-o", str(out), "-j", "2"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout == ""
    assert sorted(out.read_text().splitlines()) == sorted(want)
    r = subprocess.run([exe, "--debug-hits", str(data), str(fa1)], capture_output=True, text=True)
    hit_lines = [l for l in r.stdout.splitlines() if l.count("\t") >= 6]
    assert hit_lines and hit_lines[0].split("\t")[0] == recs[0][1][:8]
    # errors: no input files -> usage, exit 1; no database -> message, exit 1; help -> exit 0
    assert subprocess.run([exe, str(data)], capture_output=True).returncode == 1
    r = subprocess.run([exe, str(tmp_path), str(fa1)], capture_output=True, text=True)
    assert r.returncode == 1 and "does not exist" in r.stderr
    assert subprocess.run([exe, "-h"], capture_output=True).returncode == 0


# ---- GPU: batch lookups and the whole chain ---------------------------------------------------------
@pytest.mark.gpu
def test_gpu_lookup_matches_host_windows_and_membership():
    """sigk_lookup on the table of a real build == for_each_kmer + dictionary membership, position by position."""
    from signature_kmers_b200.builder import GpuSignatureBuilder

    seqs, funcs = random_proteins(41, n_families=30, members=(3, 10), length=(30, 300), sub_rate=0.06, n_functions=15)
    b = GpuSignatureBuilder(device=0)
    b.set_proteins(pack(seqs, funcs))
    table = b.build()
    index = {k: i for i, k in enumerate(table.kmer_strings())}
    rng = random.Random(3)
    queries = []
    for s in seqs[:60]:
        s = s.decode("latin-1") if isinstance(s, bytes) else s
        queries.append("".join(c if rng.random() > 0.02 else rng.choice("X*bz" + AA) for c in s))
    queries += ["", "ACDEF", "X" * 20, "ACDEFGHI", "ACDEFGHIX"]
    starts = np.zeros(len(queries) + 1, dtype=np.uint64)
    starts[1:] = np.cumsum([len(q) for q in queries])
    res = np.frombuffer("".join(queries).encode("latin-1"), dtype=np.uint8)
    rows = b.lookup(res, starts)
    want = np.full(len(res), 0xFFFFFFFF, dtype=np.uint32)
    for q, s0 in zip(queries, starts[:-1]):
        for kmer, off in co.for_each_kmer(q):
            if kmer in index:
                want[int(s0) + off] = index[kmer]
    np.testing.assert_array_equal(rows, want)
    assert (rows != 0xFFFFFFFF).sum() > 1000
    # a table given from outside: same answers
    b2 = GpuSignatureBuilder(device=0)
    b2.set_table(table.kmer)
    np.testing.assert_array_equal(b2.lookup(res, starts), want)
    b.close(); b2.close()


@pytest.mark.gpu
def test_chain_build_store_call(tmp_path):
    """config 5 in small: kmers-build-signatures (GPU) with --perfect-hash -> kmer_data.sigk + recall.report.d,
    then kmers-call-functions (host lookups and --gpu lookups) on the training FASTA; calls equal the Python
    oracle run on the table the CPU oracle of the build produces from the same packed proteins."""
    from oracle import oracle_c
    from signature_kmers_b200.capi import PackedProteins
    from signature_kmers_b200.synth import Synth
    from tests.test_host_dropin import read_packed

    subprocess.run(["make", "-C", os.path.join(PKG, "host")], check=True, capture_output=True)
    s = Synth(n_proteins=1500, n_functions=60, n_genomes=5, seed=17)
    tree = tmp_path / "tree"
    s.write_tree(str(tree))
    out = tmp_path / "out"
    cmd = [os.path.join(PKG, "kmers-build-signatures"), "-D", str(tree / "Annotations" / "0"), "-F", str(tree / "Seqs"),
           "--kmer-data-dir", str(out), "--final-kmers", "final.kmers", "--perfect-hash", "kmer_data.mph", "--sorted-files", "--n-threads", "3"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    dump = tmp_path / "packed.bin"
    subprocess.run(cmd + ["--dump-packed", str(dump)], check=True, capture_output=True)
    res, starts, func, sid = read_packed(dump)
    table, _ = oracle_c.oracle_build(PackedProteins(res.copy(), starts.copy(), func.copy(), sid.copy()))
    rows = {k: (int(table.avg_from_end[i]), int(table.function_index[i]), int(table.mean[i]), int(table.median[i]), int(table.var[i]))
            for i, k in enumerate(table.kmer_strings())}
    names = {}
    for line in open(out / "function.index").read().splitlines():
        idx, name = line.split("\t")[:2]
        names[int(idx)] = name
    names = [names.get(i, "") for i in range(max(names) + 1)]
    fc = co.FunctionCaller(rows, names)

    fasta_files = sorted((tree / "Seqs").iterdir())
    want = []
    for f in fasta_files:
        recs = [(blk.split("\n", 1)[0].split()[0], "".join(blk.split("\n")[1:])) for blk in f.read_text().split(">")[1:]]
        for rid, seq in recs:
            fi, fn, score = fc.call(seq)
            want.append("%s\t%s\t%d\t%s" % (rid, fn, fi, co.format_score(score)))
    exe = os.path.join(PKG, "kmers-call-functions")
    host_run = subprocess.run([exe, str(out)] + [str(f) for f in fasta_files], capture_output=True, text=True)
    assert host_run.returncode == 0, host_run.stderr
    assert host_run.stdout.splitlines() == want
    gpu_run = subprocess.run([exe, "--gpu", "0", str(out)] + [str(f) for f in fasta_files], capture_output=True, text=True)
    assert gpu_run.returncode == 0, gpu_run.stderr
    assert gpu_run.stdout == host_run.stdout
    called = [l for l in want if not l.endswith("\t65535\t0")]
    assert len(called) > 0.8 * len(want)                # the training set is recalled
    # recall.report.d (GPU lookups) == the --host-recall variant; every reported row is a call that differs
    rep = {f.name: f.read_text() for f in (out / "recall.report.d").iterdir()}
    assert set(rep) == {f.name for f in fasta_files}
    out2 = tmp_path / "out2"
    r2 = subprocess.run([c if c != str(out) else str(out2) for c in cmd] + ["--host-recall"], capture_output=True, text=True)
    assert r2.returncode == 0, r2.stderr
    assert rep == {f.name: f.read_text() for f in (out2 / "recall.report.d").iterdir()}
    calls_by_id = {l.split("\t")[0]: l.split("\t") for l in want}
    n_rows = 0
    for text in rep.values():
        for line in text.splitlines():
            rid, old, old_stripped, new, idx, score = line.split("\t")
            assert calls_by_id[rid][1:] == [new, idx, score] and old_stripped != new
            n_rows += 1
    # ... and every call that differs from the assignment is reported
    assigned = {}
    for f in (tree / "Annotations" / "0").iterdir():
        for line in f.read_text().splitlines():
            rid, fn = line.split("\t")[:2]
            assigned[rid] = fn
    assert n_rows == sum(1 for rid, c in calls_by_id.items() if assigned.get(rid, "") != c[1])
