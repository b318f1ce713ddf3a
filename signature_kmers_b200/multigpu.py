"""One process per GPU: launcher-side plumbing for the range-partitioned build.

The data path (sampled splitters, encode + route into the owners' landing zones
over NVLink, per-rank sort/reduce) lives in csrc/comm.cu behind sigk_comm_join + sigk_build;
torch.distributed is used here only to ship the 128-byte communicator id,
for barriers, and to combine timings.  Test/bench driver, not the product.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import time

import numpy as np

from . import capi
from .capi import KeptTable, PackedProteins


def rank_slice(n: int, rank: int, world: int):
    """Contiguous chunk `rank` of n canonical positions (chunks in rank order = canonical order)."""
    return n * rank // world, n * (rank + 1) // world


def bind_to_gpu_numa_node(device: int):
    """Pin this process to the CPUs next to its GPU (sysfs local_cpulist of the GPU's PCI function) before any
    pinned host buffer exists: cudaMallocHost places pages near the calling thread, and with one process per GPU
    and no binding every rank's result buffers tend to land on one socket — the device-to-host copies of the other
    socket's GPUs then cross the inter-socket link (round-1 8-GPU run: 86 GB/s for all eight downloads together).
    Returns the CPU list it bound to, or None when the topology cannot be read (then nothing changes)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = int(visible.split(",")[device]) if visible and all(x.strip().isdigit() for x in visible.split(",")) else device
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        with open(f"/sys/bus/pci/devices/{int(dom, 16):04x}:{rest.lower()}/local_cpulist") as f:
            text = f.read().strip()
        cpus = set()
        for part in text.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


def init_process_group():
    import torch
    import torch.distributed as dist

    if not dist.is_initialized():
        backend = "cpu:gloo,cuda:nccl" if torch.cuda.is_available() else "gloo"
        dist.init_process_group(backend=backend)
    return dist


def join_communicator(builder, rank: int, world: int):
    """Rank 0 makes the NCCL id, everybody joins (sigk_comm_make_id / sigk_comm_join)."""
    import torch.distributed as dist

    lib = capi.load_library()
    ident = [None]
    if rank == 0:
        buf = C.create_string_buffer(capi.SIGK_COMM_ID_BYTES)
        rc = lib.sigk_comm_make_id(buf)
        if rc != 0:
            raise capi.SigkError(f"sigk_comm_make_id failed ({rc}): {lib.sigk_last_error(None).decode()}")
        ident[0] = buf.raw
    dist.broadcast_object_list(ident, src=0)
    keep = C.create_string_buffer(ident[0], capi.SIGK_COMM_ID_BYTES)
    rc = lib.sigk_comm_join(builder.h, keep)
    if rc != 0:
        raise capi.SigkError(f"sigk_comm_join failed ({rc}): {lib.sigk_last_error(builder.h).decode()}")


def concat_tables(tables) -> KeptTable:
    """Per-rank slices in rank order -> the whole kept table (counters are already job-wide).  Every slice is in
    table order (include/sigk.h): its rows without a lower-case residue first.  K-mer ranges are cut on the
    case-folded k-mer and ascend with the rank, so the first sections in rank order followed by the second
    sections in rank order are the whole table in table order."""
    t0 = tables[0]
    ups = [t.n_upper for t in tables]

    def cat(name):
        return np.concatenate([getattr(t, name)[:u] for t, u in zip(tables, ups)] + [getattr(t, name)[u:] for t, u in zip(tables, ups)])

    return KeptTable(
        kmer=cat("kmer").reshape(-1, 8),
        avg_from_end=cat("avg_from_end"), function_index=cat("function_index"), mean=cat("mean"),
        median=cat("median"), var=cat("var"),
        n_occurrences=t0.n_occurrences, n_distinct_kmers=t0.n_distinct_kmers,
        distinct_signatures=t0.distinct_signatures, num_seqs_with_a_signature=t0.num_seqs_with_a_signature,
        distinct_functions=t0.distinct_functions, seqs_with_func=t0.seqs_with_func, n_upper=sum(ups),
    )


def weak_scaling_params(workload: str, world: int) -> dict:
    """Per-GPU work fixed at the single-GPU workload: world x the proteins, genomes and functions.  Past 65 535 kept
    functions the generator's index assignment wraps exactly like the reference's `unsigned short next`
    (src/function_map.h:324-330): later functions alias earlier indices and index 0xFFFF proteins are skipped
    (src/signature_build.tcc:155-158) — the config-3 situation (8 x 20 000 = 160 000 functions at 8 GPUs)."""
    from .synth import CONFIGS

    kw = dict(CONFIGS[workload])
    kw["n_functions"] = kw["n_functions"] * world
    kw["n_proteins"] = kw["n_proteins"] * world
    kw["n_genomes"] = kw["n_genomes"] * world
    return kw


def run_bench(args, rank, world, local_rank, metric, unit):
    import torch
    import torch.distributed as dist

    from .builder import GpuSignatureBuilder
    from .synth import Synth

    dist = init_process_group()
    torch.cuda.set_device(local_rank)
    # weak scaling (default): config x world, per-GPU work fixed.  --scaling strong: the named configuration as it is,
    # split over the ranks (config 3 = 20 M proteins / 100 K functions is the one BASELINE.json names for 2/4/8 GPUs)
    strong = getattr(args, "scaling", "weak") == "strong"
    if strong:
        from .synth import CONFIGS

        kw = dict(CONFIGS[args.workload])
    else:
        kw = weak_scaling_params(args.workload, world)
    synth = Synth(**kw)
    lo, hi = rank_slice(synth.n_proteins, rank, world)
    bound = bind_to_gpu_numa_node(local_rank)
    builder = GpuSignatureBuilder(device=local_rank, rank=rank, world=world)
    join_communicator(builder, rank, world)
    proteins = synth.packed(lo, hi, out_alloc=builder.host_alloc)
    builder.set_proteins(proteins)
    builder.upload()
    for _ in range(args.warmup):
        builder.build_device()
    builder.synchronize()
    dist.barrier()

    sampler = None
    if rank == 0:
        from bench import ClockSampler  # the launcher script

        sampler = ClockSampler(local_rank)
        sampler.start()
        time.sleep(0.3)
    dist.barrier()
    torch.cuda.synchronize()
    builder.event_record(0)
    for _ in range(args.steps):
        builder.build_device()
    builder.event_record(1)
    builder.synchronize()
    dist.barrier()
    dev_ms = builder.event_elapsed_ms(0, 1)
    builder.download()
    counts = builder.result_counts()
    tm = builder.timings()

    # end to end: pinned host arrays in, this rank's slice of the kept table back in host memory
    builder.build(fetch=False)
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        builder.build(fetch=False)
    dist.barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps

    stats = torch.tensor([dev_ms, e2e_s, float(counts["n_kept"]), float(proteins.residues.nbytes + proteins.starts.nbytes
                          + proteins.function_index.nbytes + proteins.seq_id.nbytes)], dtype=torch.float64)
    mx = stats.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = stats.clone()
    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    clk = sampler.stop() if sampler else None
    stage_keys = ("encode_ms", "exchange_ms", "histogram_ms", "sort_ms", "side_sort_ms", "reduce_ms", "reduce_count_ms", "reduce_emit_ms",
                  "reduce_groups_ms", "reduce_comm_ms", "order_stats_ms", "squeeze_ms", "device_total_ms")
    # every rank's stage times of its last build: a stage that holds a collective absorbs the other ranks' lateness,
    # so rank 0's column alone cannot say which stage is slow
    st = torch.tensor([float(tm[k]) for k in stage_keys] + [float(tm["records_sorted"])], dtype=torch.float64)
    all_st = [torch.zeros_like(st) for _ in range(world)]
    dist.all_gather(all_st, st)
    if rank == 0:
        occ = counts["n_occurrences"]          # job-wide after the statistics reduction
        ms_per_step = float(mx[0]) / args.steps
        passes = int(tm["sort_passes"])
        line = {
            "metric": metric, "value": occ / (ms_per_step * 1e-3), "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": (f"{args.workload} over {world} GPUs (strong scaling: " if strong else f"{args.workload} x {world} (weak scaling: ")
                                   + f"{kw['n_proteins']} proteins, {kw['n_functions']} functions, {kw['n_genomes']} genomes; rank r encodes canonical chunk r)",
                       "kept_functions": synth.kept_functions,
                       "occurrences_per_step": occ, "distinct_kmers": counts["n_distinct_kmers"],
                       "partition": "k-mer code ranges by sampled splitters; the encode kernel stores each 12-byte record into its owner GPU's landing zone over NVLink (CUDA IPC), NCCL for the small collectives",
                       "K": 8, "record_bytes": 12, "sort_passes": passes,
                       "l2": "inputs larger than L2", "timed": "max over ranks of CUDA-event time on the library stream"},
            "clocks": clk,
            "e2e": {"value": occ / float(mx[1]), "unit": unit, "h2d_bytes_per_step": int(sm[3]),
                    "d2h_bytes_per_step": int(sm[2]) * 18, "ms_per_step": 1e3 * float(mx[1]),
                    "api": "sigk_build per rank (C ABI, pinned host buffers)",
                    "host_binding": f"each rank bound to its GPU's local CPUs (rank 0: {len(bound)} CPUs)" if bound else "none (topology unreadable)"},
            "gpu_launches": int(tm["kernel_launches"]) * args.steps * world,
            "roofline": None,
            "rank0_stage_ms": {k: tm[k] for k in stage_keys},
            "stage_ms_min_over_ranks": {k: min(float(t[i]) for t in all_st) for i, k in enumerate(stage_keys)},
            "stage_ms_max_over_ranks": {k: max(float(t[i]) for t in all_st) for i, k in enumerate(stage_keys)},
            "records_sorted_per_rank": [int(t[len(stage_keys)]) for t in all_st],
            "cpu_baseline": None,
        }
        pass_ms = list(tm["pass_ms"][:passes])
        if pass_ms and tm["sort_ms"] > 0:
            from bench import measured_peak_gbs

            peak, src = measured_peak_gbs()
            # what rank 0's pass kernel moved: the records of rank 0's k-mer range (the library reports the count)
            records0 = int(tm["records_sorted"])
            plain = pass_ms[1:] if len(pass_ms) > 1 else pass_ms
            launch_ms = sum(plain) / len(plain)
            alg_bytes = 24 * records0
            achieved = alg_bytes / (launch_ms * 1e-3) / 1e9
            line["roofline"] = {"bound": "hbm", "kernel": "onesweep_pass_kernel (rank 0)", "peak": peak, "unit": "GB/s",
                                "peak_source": src, "pass_ms": pass_ms, "launch_ms": launch_ms,
                                "algorithmic_bytes_per_launch": alg_bytes, "records": records0,
                                "achieved": achieved, "frac": achieved / peak, "traffic": None}
            # the exchange: what rank 0's encode + route kernel stored into the other GPUs' landing zones over NVLink
            if tm["encode_ms"] > 0 and tm["exchange_bytes_out"]:
                out_gbs = tm["exchange_bytes_out"] / (tm["encode_ms"] * 1e-3) / 1e9
                line["nvlink"] = {"rank0_bytes_out_per_step": int(tm["exchange_bytes_out"]), "encode_route_ms": tm["encode_ms"],
                                  "achieved_gbs_out": out_gbs, "peak_gbs_per_direction": 900.0, "frac": out_gbs / 900.0}
        from bench import emit_json

        emit_json(line)
    builder.close()
    dist.barrier()
    dist.destroy_process_group()
