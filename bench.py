#!/usr/bin/env python
"""bench.py — k-mer occurrences/s into the kept-signature table on B200.

A step is one pass of the signature-generation hot path (window count -> encode
fused with the first radix pass -> onesweep passes -> segment reduce ->
keep/compact; the reference's extract_kmers + process_kmers,
src/signature_build.tcc:47-293) over one synthetic protein set.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config2]
  python bench.py --impl reference ...     # the CPU path, timed on the host cores

`value` is timed on the device (CUDA events on the library's stream) with the
packed proteins already resident in HBM; `e2e` is the same metric through the
C-ABI call sigk_build with pinned HOST buffers in and the kept table back in
host memory (H2D and D2H inside the timed region).  One JSON line on stdout.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

if __name__ == "__main__":
    sys.modules.setdefault("bench", sys.modules["__main__"])   # multigpu.py imports helpers from this script

METRIC = "kmer occurrences/s into kept-signature table"
UNIT = "occurrences/s"
RECORD_BYTES = 12
KEY_BYTES = 8


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# stdout carries exactly one JSON line: everything else that writes to fd 1 (NCCL's version banner, library
# chatter) is sent to stderr, and emit_json() writes to the saved descriptor.
_REAL_STDOUT = None


def capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit_json(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="sigk_clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1])); mx.append(float(c[2]))
                except ValueError:
                    continue
                for nm, v in zip(names, c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            # the GPU idles between phases of the bench; the clock under load is the upper half
            hi = sorted(sm)[len(sm) // 2:]
            out.update(sm_mhz=statistics.median(hi), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def pinned_copy(builder, a: np.ndarray) -> np.ndarray:
    buf = builder.host_alloc(a.nbytes)
    out = buf.view(a.dtype)[: a.size].reshape(a.shape)
    out[...] = a
    return out


def algorithmic_bytes_per_occurrence(passes: int, fused: bool) -> dict:
    """B_alg = 1 + Kb + (2P + 2) R with this build's R, Kb, P — only the passes executed and the
    one histogram read.  Fused (one GPU): the histogram is a second read of the residues (1 B instead of Kb) and
    the first pass reads residues instead of records, so the encode's write is the first pass's write:
    1 (count) + 1 + R (encode + first pass) + 2 R (P - 1) + R (reduce)."""
    if fused:
        enc = 1 + RECORD_BYTES
        hist = 1
        sort = 2 * RECORD_BYTES * (passes - 1)
    else:
        enc = 1 + RECORD_BYTES
        hist = KEY_BYTES
        sort = 2 * RECORD_BYTES * passes
    red = RECORD_BYTES
    return dict(encode_and_first_pass=enc, histogram=hist, sort=sort, reduce=red, total=enc + hist + sort + red)


def cpu_baseline(proteins, sample_proteins: int, threads: int):
    """The oracle port of the reference's CPU path, all host threads, on a bounded sample."""
    from oracle import oracle_c

    n = min(sample_proteins, proteins.n_proteins)
    sample = proteins.slice(0, n)
    t, secs = oracle_c.oracle_build(sample, n_threads=threads, flags=oracle_c.NO_SORT, want_table=False)
    return dict(value=t.n_occurrences / secs, unit=UNIT, cores=threads, kind="port",
                sample=f"first {n} proteins of the workload in canonical order ({t.n_occurrences} occurrences, {secs:.2f} s of extract+process)")


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  The reference
    itself cannot be built in this image (no Boost/TBB/cmph), so this times the oracle
    port (kind "port") with all host threads on a bounded sample of the same workload."""
    if rank != 0:
        return
    from oracle import oracle_c
    from signature_kmers_b200.synth import Synth

    threads = os.cpu_count() or 1
    synth = Synth.config(args.workload)
    n = min(args.cpu_sample_proteins, synth.n_proteins)
    sample = synth.packed(0, n)
    times, occ = [], 0
    for i in range(args.warmup + args.steps):
        t, secs = oracle_c.oracle_build(sample, n_threads=threads, flags=oracle_c.NO_SORT, want_table=False)
        occ = t.n_occurrences
        if i >= args.warmup:
            times.append(secs)
    total = sum(times)
    value = occ * len(times) / total
    desc = f"first {sample.n_proteins} gated proteins of {args.workload} ({occ} occurrences per step)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": args.workload, "sample": desc, "timed": "extract_kmers + process_kmers on host cores"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_json(line)


def run_gpu(args, rank, world, local_rank):
    from signature_kmers_b200.builder import GpuSignatureBuilder
    from signature_kmers_b200.synth import CONFIGS, Synth

    if world > 1:
        from signature_kmers_b200 import multigpu      # noqa: F401  (one process per GPU; see multigpu.py)
        return multigpu.run_bench(args, rank, world, local_rank, METRIC, UNIT)

    t0 = time.time()
    # host buffers next to the GPU (see multigpu.bind_to_gpu_numa_node); the CPU baseline below gets all cores back
    from signature_kmers_b200.multigpu import bind_to_gpu_numa_node

    all_cpus = os.sched_getaffinity(0)
    bound = bind_to_gpu_numa_node(local_rank)
    builder = GpuSignatureBuilder(device=local_rank)
    synth = Synth.config(args.workload)
    proteins = synth.packed(out_alloc=builder.host_alloc)
    proteins.starts = pinned_copy(builder, proteins.starts)
    proteins.function_index = pinned_copy(builder, proteins.function_index)
    proteins.seq_id = pinned_copy(builder, proteins.seq_id)
    log(f"[bench] generated {args.workload}: {proteins.n_proteins} proteins, {len(proteins.residues)} residues in {time.time() - t0:.1f}s")

    builder.set_proteins(proteins)
    builder.upload()
    for _ in range(args.warmup):
        builder.build_device()
    builder.synchronize()

    clocks = ClockSampler(local_rank)
    clocks.start()
    time.sleep(0.3)

    # ---- device-timed region: K steps, inputs resident in HBM
    builder.event_record(0)
    for _ in range(args.steps):
        builder.build_device()
    builder.event_record(1)
    builder.synchronize()
    dev_ms = builder.event_elapsed_ms(0, 1)
    builder.download()
    counts = builder.result_counts()
    tm = builder.timings()
    occ = counts["n_occurrences"]
    ms_per_step = dev_ms / args.steps
    value = occ / (ms_per_step * 1e-3)

    # ---- end to end through the C ABI: pinned host arrays in, kept table in host memory out
    builder.build(fetch=False)            # sizes the pinned result buffers once
    e2e_t0 = time.perf_counter()
    for _ in range(args.steps):
        builder.build(fetch=False)        # the table lands in pinned host memory owned by the handle
    e2e_s = (time.perf_counter() - e2e_t0) / args.steps
    h2d = proteins.residues.nbytes + proteins.starts.nbytes + proteins.function_index.nbytes + proteins.seq_id.nbytes
    d2h = counts["n_kept"] * 18 + 2 * 65536 * 4 + 96
    clk = clocks.stop()
    tm_e2e = builder.timings()            # copies of the last end-to-end build: steady state (pinned buffers already sized)

    passes = int(tm["sort_passes"])
    fused = os.environ.get("SIGK_NO_FUSED") is None
    pass_ms = [x for x in tm["pass_ms"][:passes]]
    # the dominant kernel: the plain onesweep passes (3 launches per build; the first pass is the fused kernel)
    plain_ms = pass_ms[1:] if fused and passes > 1 else pass_ms
    peak, peak_src = measured_peak_gbs()
    alg_bytes_per_launch = 2 * RECORD_BYTES * occ
    mean_pass_ms = sum(plain_ms) / max(1, len(plain_ms))
    achieved = alg_bytes_per_launch / (mean_pass_ms * 1e-3) / 1e9 if mean_pass_ms > 0 else 0.0
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "onesweep_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("workload") == args.workload:
            traffic = tj.get("dram_bytes_per_launch")
            traffic_src = "constant from profiles/ (%s): ncu --set full of this command, not this run" % tj.get("source", "onesweep_traffic.json")
    except Exception:
        pass
    balg = algorithmic_bytes_per_occurrence(passes, fused)

    cpu = None
    os.sched_setaffinity(0, all_cpus)
    if not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline(proteins, args.cpu_sample_proteins, os.cpu_count() or 1)
        except Exception as e:  # the oracle is only a reported baseline
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {
            "workload": args.workload, **{k: v for k, v in CONFIGS[args.workload].items()},
            "occurrences_per_step": occ, "distinct_kmers": counts["n_distinct_kmers"], "kept_kmers": counts["n_kept"],
            "K": 8, "record_bytes": RECORD_BYTES, "sort_passes": passes, "first_pass_fused_with_encode": fused,
            "l2": "inputs larger than L2 (residues %.0f MB, records %.1f GB per step)" % (proteins.residues.nbytes / 1e6, occ * RECORD_BYTES / 1e9),
            "timed": "window count + encode/first pass + onesweep passes + segment reduce + keep/compact (CUDA events on the library stream)",
        },
        "clocks": clk,
        "e2e": {"value": occ / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1e3 * e2e_s, "api": "sigk_build (C ABI, pinned host buffers)",
                "host_binding": f"process bound to the GPU's {len(bound)} local CPUs while the pinned buffers are allocated" if bound else "none"},
        "gpu_launches": int(tm["kernel_launches"]) * args.steps,
        "roofline": {"bound": "hbm", "kernel": "onesweep_pass_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if peak else None, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes_per_launch, "launch_ms": mean_pass_ms, "launches_per_step": len(plain_ms),
                     "pass_ms": pass_ms,
                     "first_pass": {"kernel": "encode_sort_kernel" if fused else "onesweep_pass_kernel<FIRST>", "ms": pass_ms[0] if pass_ms else None,
                                    "algorithmic_bytes_per_launch": (1 + RECORD_BYTES if fused else 2 * RECORD_BYTES) * occ,
                                    "achieved": ((1 + RECORD_BYTES if fused else 2 * RECORD_BYTES) * occ / (pass_ms[0] * 1e-3) / 1e9) if pass_ms and pass_ms[0] > 0 else None}},
        "pipeline": {"b_alg_per_occurrence": balg, "achieved_gbs": balg["total"] * occ / (ms_per_step * 1e-3) / 1e9,
                     "frac_of_peak": balg["total"] * occ / (ms_per_step * 1e-3) / 1e9 / peak,
                     "stage_ms": {**{k: tm[k] for k in ("encode_ms", "count_ms", "histogram_ms", "sort_ms", "side_sort_ms", "reduce_ms", "reduce_count_ms", "reduce_emit_ms",
                                                          "reduce_groups_ms", "order_stats_ms", "squeeze_ms", "device_total_ms")},
                                  "h2d_ms": tm_e2e["h2d_ms"], "d2h_ms": tm_e2e["d2h_ms"]}},
        "cpu_baseline": cpu,
    }
    emit_json(line)
    builder.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="sigk", choices=["sigk", "reference"])
    ap.add_argument("--workload", default="config2", choices=["config1", "config2", "config3", "config4"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="multi-GPU: weak = workload x N (default, the driver's scaling run); strong = the named workload split over N GPUs")
    ap.add_argument("--cpu-sample-proteins", type=int, default=600_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "sigk":
        log("[bench] note: fewer than 3 warm-up steps")

    capture_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
