#!/usr/bin/env python
"""Attribute an ncu SASS-level source page to CUDA source lines.

usage: ncu_lines.py REPORT.ncu-rep KERNEL_REGEX CUBIN_NAME(e.g. reduce) [top]

ncu's CSV source page lists SASS with per-instruction counters but no line
numbers; nvdisasm -g on the cubin of the same build lists the same SASS with
line info.  The two listings are joined by instruction order.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_lines(cubin_name, kernel_regex, want_len=None):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "signature_kmers_b200", "libsigk.so")], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.startswith(cubin_name + ".")][0]
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    cur_fn, cur_line, res = None, None, {}
    for ln in out.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            cur_fn = m.group(1); res.setdefault(cur_fn, []); continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and cur_fn:
            res[cur_fn].append((cur_line, m.group(2).strip()))
    cands = [v for fn, v in res.items() if re.search(kernel_regex, fn)]
    if not cands:
        raise SystemExit("kernel not found in cubin")
    if want_len is None:
        return cands[0]
    # several template instantiations can match: take the one whose SASS length fits the report
    return min(cands, key=lambda v: abs(len(v) - want_len))


def main():
    rep, kre, cubin = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    hdr = rows[hi]
    ii, si, sa = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    body = []
    for r in rows[hi + 1:]:
        if len(r) <= ii or r[0] == "Kernel Name":
            if body:
                break       # first launch only
            continue
        try:
            body.append((r[si].strip(), int(r[ii]), int(r[sa])))
        except ValueError:
            pass
    sass = sass_lines(cubin, kre, len(body))
    if len(sass) != len(body):
        print(f"warning: {len(sass)} SASS instructions in cubin vs {len(body)} in report", file=sys.stderr)
    agg = {}
    for (line, _), (_, n, smp) in zip(sass, body):
        a = agg.setdefault(line, [0, 0]); a[0] += n; a[1] += smp
    tot = sum(a[0] for a in agg.values()); tots = sum(a[1] for a in agg.values()) or 1
    src = {}
    print(f"total warp instructions {tot/1e9:.2f} G, samples {tots}")
    for line, (n, smp) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        if line:
            f = os.path.join(ROOT, "signature_kmers_b200", "csrc", line[0])
            if f not in src and os.path.exists(f):
                src[f] = open(f).read().splitlines()
            if f in src and line[1] <= len(src[f]):
                text = src[f][line[1] - 1].strip()[:100]
        print(f"{n/1e6:9.1f}M {100*n/tot:5.1f}%  stall {100*smp/tots:5.1f}%  {line}  {text}")


if __name__ == "__main__":
    main()
