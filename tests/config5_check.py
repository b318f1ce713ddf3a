"""Config 5 (BASELINE.json configs[4]) at size: GPU kept table -> table file -> kmers-call-functions on 100 K queries.

  python tests/config5_check.py [--proteins 200000] [--functions 2000] [--queries 100000] [--oracle-sample 1500]

Builds a synthetic tree, runs this repo's kmers-build-signatures (GPU build, --perfect-hash -> kmer_data.sigk, recall
pass through sigk_lookup), then kmers-call-functions on the first --queries training proteins twice (host lookups,
--gpu lookups): the two outputs must be identical, and a sample of the calls must equal the Python oracle
(oracle/call_oracle.py) evaluated on the table file.  Prints timings and one JSON line.  Needs a GPU; not part of pytest.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "signature_kmers_b200")

from oracle import call_oracle as co  # noqa: E402
from signature_kmers_b200.synth import Synth  # noqa: E402


class FileTable:
    """kmer_data.sigk as a mapping kmer -> (avg_from_end, function_index, mean, median, var) for the oracle."""

    def __init__(self, path):
        raw = np.memmap(path, dtype=np.uint8, mode="r")
        assert bytes(raw[:8]) == b"SIGKTBL1"
        n = int(np.frombuffer(raw[8:16], dtype=np.uint64)[0])
        self.n = n
        # big-endian view: integer order = byte order; native uint64 copy so that searchsorted compares integers
        self.keys = np.frombuffer(raw[16:16 + 8 * n], dtype=">u8").astype(np.uint64)
        off = 16 + 8 * n
        self.cols = [np.frombuffer(raw[off + 2 * n * c: off + 2 * n * (c + 1)], dtype=np.uint16) for c in range(5)]
        # table order (include/sigk.h): rows without a lower-case residue first, in byte order; the (few) others after
        # them, kept here in a dictionary
        lower = (self.keys & np.uint64(0x2020202020202020)) != 0
        self.n_upper = int(np.argmax(lower)) if lower.any() else n
        assert not lower[:self.n_upper].any() and lower[self.n_upper:].all(), "two sections"
        assert (np.diff(self.keys[:self.n_upper].astype(np.int64)) > 0).all() if self.n_upper > 1 else True
        self.side = {int(k): self.n_upper + i for i, k in enumerate(self.keys[self.n_upper:])}

    def get(self, kmer):
        key = int.from_bytes(kmer.encode("latin-1"), "big")
        if key & 0x2020202020202020:
            i = self.side.get(key)
            return None if i is None else tuple(int(c[i]) for c in self.cols)
        i = int(np.searchsorted(self.keys[:self.n_upper], np.uint64(key)))
        if i < self.n_upper and int(self.keys[i]) == key:
            return tuple(int(c[i]) for c in self.cols)
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--proteins", type=int, default=200_000)
    ap.add_argument("--functions", type=int, default=2_000)
    ap.add_argument("--genomes", type=int, default=10)
    ap.add_argument("--queries", type=int, default=100_000)
    ap.add_argument("--oracle-sample", type=int, default=1500)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    args = ap.parse_args()
    subprocess.run(["make", "-C", os.path.join(PKG, "host")], check=True, capture_output=True)
    with tempfile.TemporaryDirectory() as tmp:
        s = Synth(n_proteins=args.proteins, n_functions=args.functions, n_genomes=args.genomes, seed=5)
        tree = os.path.join(tmp, "tree")
        s.write_tree(tree)
        out = os.path.join(tmp, "out")
        t0 = time.perf_counter()
        r = subprocess.run([os.path.join(PKG, "kmers-build-signatures"), "-D", os.path.join(tree, "Annotations", "0"), "-F", os.path.join(tree, "Seqs"),
                            "--kmer-data-dir", out, "--final-kmers", "final.kmers", "--perfect-hash", "kmer_data.mph", "--sorted-files",
                            "--n-threads", str(args.threads)], capture_output=True, text=True)
        t_build = time.perf_counter() - t0
        assert r.returncode == 0, r.stderr[-2000:]
        kept = [l for l in r.stdout.splitlines() if l.startswith("Kept ")][0]
        # the query set: the first --queries records of the training files, written to a few query files
        records = []
        for f in sorted(os.listdir(os.path.join(tree, "Seqs"))):
            for blk in open(os.path.join(tree, "Seqs", f)).read().split(">")[1:]:
                head, _, body = blk.partition("\n")
                records.append((head.split()[0], body.replace("\n", "")))
                if len(records) >= args.queries:
                    break
            if len(records) >= args.queries:
                break
        qfiles = []
        per = (len(records) + 7) // 8
        for i in range(0, len(records), per):
            p = os.path.join(tmp, "q%d.fa" % (i // per))
            with open(p, "w") as fh:
                for rid, seq in records[i:i + per]:
                    fh.write(">%s\n%s\n" % (rid, seq))
            qfiles.append(p)
        windows = sum(max(0, len(seq) - 7) for _, seq in records)
        exe = os.path.join(PKG, "kmers-call-functions")
        t0 = time.perf_counter()
        host = subprocess.run([exe, "-j", str(args.threads), out] + qfiles, capture_output=True, text=True)
        t_host = time.perf_counter() - t0
        assert host.returncode == 0, host.stderr[-2000:]
        t0 = time.perf_counter()
        gpu = subprocess.run([exe, "-j", str(args.threads), "--gpu", "0", out] + qfiles, capture_output=True, text=True)
        t_gpu = time.perf_counter() - t0
        assert gpu.returncode == 0, gpu.stderr[-2000:]
        host_lines, gpu_lines = sorted(host.stdout.splitlines()), sorted(gpu.stdout.splitlines())
        assert len(host_lines) == len(records)
        assert host_lines == gpu_lines, "host and GPU lookups give different calls"
        # oracle on a sample
        names = {}
        for line in open(os.path.join(out, "function.index")).read().splitlines():
            idx, name = line.split("\t")[:2]
            names[int(idx)] = name
        names = [names.get(i, "") for i in range(max(names) + 1)]
        fc = co.FunctionCaller(FileTable(os.path.join(out, "kmer_data.sigk")), names)
        by_id = {l.split("\t")[0]: l for l in host_lines}
        step = max(1, len(records) // args.oracle_sample)
        checked = 0
        for rid, seq in records[::step]:
            fi, fn, score = fc.call(seq)
            assert by_id[rid] == "%s\t%s\t%d\t%s" % (rid, fn, fi, co.format_score(score)), (rid, by_id[rid])
            checked += 1
        called = sum(1 for l in host_lines if not l.endswith("\t65535\t0"))
        print(json.dumps({"config": "5: calls on %d queries (%d windows) against the table of %d proteins / %d functions" %
                          (len(records), windows, args.proteins, args.functions), "table": kept, "called": called,
                          "build_cli_s": round(t_build, 2), "call_host_lookups_s": round(t_host, 2), "call_gpu_lookups_s": round(t_gpu, 2),
                          "threads": args.threads, "oracle_checked": checked, "host_equals_gpu": True}))
        print("CONFIG5_CHECK_PASSED")


if __name__ == "__main__":
    main()
