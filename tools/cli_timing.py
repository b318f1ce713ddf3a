#!/usr/bin/env python
"""Where the wall time of a whole kmers-build-signatures run goes (--timings), with the host FASTA reader and with
--gpu-fasta, on a synthetic tree (default: config 5's build — 200 K proteins / 2 K functions / 20 genomes).
Prints one JSON object.  python tools/cli_timing.py [--proteins N --functions F --genomes G --threads T]"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--proteins", type=int, default=200_000)
    ap.add_argument("--functions", type=int, default=2_000)
    ap.add_argument("--genomes", type=int, default=20)
    ap.add_argument("--threads", type=int, default=16)
    args = ap.parse_args()
    from signature_kmers_b200.synth import Synth

    pkg = os.path.join(ROOT, "signature_kmers_b200")
    with tempfile.TemporaryDirectory() as tmp:
        tree = os.path.join(tmp, "tree")
        Synth(n_proteins=args.proteins, n_functions=args.functions, n_genomes=args.genomes, seed=5).write_tree(tree)
        fasta_bytes = sum(os.path.getsize(os.path.join(tree, "Seqs", f)) for f in os.listdir(os.path.join(tree, "Seqs")))
        out = {"proteins": args.proteins, "functions": args.functions, "genomes": args.genomes, "threads": args.threads, "fasta_bytes": fasta_bytes, "runs": {}}
        outputs = {}
        for name, flags in (("host_reader", []), ("gpu_fasta", ["--gpu-fasta"])):
            best = None
            for rep in range(2):                     # the second run has the files in the page cache and the GPU context warm
                dest = os.path.join(tmp, "out_%s_%d" % (name, rep))
                cmd = [os.path.join(pkg, "kmers-build-signatures"), "-D", os.path.join(tree, "Annotations", "0"), "-F", os.path.join(tree, "Seqs"),
                       "--kmer-data-dir", dest, "--final-kmers", "final.kmers", "--sigk-table", "kmer_data.sigk", "--sorted-files",
                       "--n-threads", str(args.threads), "--timings"] + flags
                t0 = time.perf_counter()
                r = subprocess.run(cmd, capture_output=True, text=True)
                wall = time.perf_counter() - t0
                if r.returncode != 0:
                    raise SystemExit(r.stderr)
                phases = {m.group(1): float(m.group(2)) for m in re.finditer(r"^\[time\] (.*): ([0-9.e+-]+) s$", r.stderr, re.M)}
                cur = {"wall_s": wall, "phases_s": phases}
                if best is None or wall < best["wall_s"]:
                    best = cur
                outputs[name] = dest
            out["runs"][name] = best
        same = True
        for fn in ("final.kmers", "kmer_data.sigk", "function.index"):
            a = open(os.path.join(outputs["host_reader"], fn), "rb").read()
            b = open(os.path.join(outputs["gpu_fasta"], fn), "rb").read()
            same = same and a == b
        out["outputs_identical"] = same
        print(json.dumps(out))


if __name__ == "__main__":
    main()
