"""Launched by torchrun (one process per GPU): the NCCL range-partitioned build
against the CPU oracle and against the single-GPU build of the same proteins.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from signature_kmers_b200 import multigpu  # noqa: E402
from signature_kmers_b200.builder import GpuSignatureBuilder  # noqa: E402
from signature_kmers_b200.synth import Synth  # noqa: E402
from tests.util import assert_tables_equal, pack, random_proteins  # noqa: E402


def build_distributed(p_all, rank, world, local_rank):
    import torch.distributed as dist

    lo, hi = multigpu.rank_slice(p_all.n_proteins, rank, world)
    say = lambda what: print(f"[rank {rank}] {what}", file=sys.stderr, flush=True)
    say(f"proteins {lo}..{hi}")
    b = GpuSignatureBuilder(device=local_rank, rank=rank, world=world)
    multigpu.join_communicator(b, rank, world)
    say("joined")
    b.set_proteins(p_all.slice(lo, hi))
    t = b.build()
    say("built")
    again = b.build()                  # a second build on the same communicator gives the same slice
    assert_tables_equal(t, again, what=f"rank {rank} rebuild")
    parts = [None] * world
    dist.all_gather_object(parts, t)
    say("gathered")
    b.close()
    say("closed")
    return multigpu.concat_tables(parts), [x.n_kept for x in parts]


def build_sequence_on_one_communicator(cases, rank, world, local_rank):
    """Several inputs through ONE handle and communicator: the landing zones grow (re-exported and re-imported
    collectively), then are reused for a smaller input."""
    import torch.distributed as dist

    b = GpuSignatureBuilder(device=local_rank, rank=rank, world=world)
    multigpu.join_communicator(b, rank, world)
    out = []
    for _, p_all in cases:
        lo, hi = multigpu.rank_slice(p_all.n_proteins, rank, world)
        b.set_proteins(p_all.slice(lo, hi))
        t = b.build()
        parts = [None] * world
        dist.all_gather_object(parts, t)
        out.append(multigpu.concat_tables(parts))
    b.close()
    return out


def main():
    import torch

    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist = multigpu.init_process_group()
    torch.cuda.set_device(local_rank)
    cases = []
    seqs, funcs = random_proteins(71, n_families=60, members=(2, 30), length=(20, 300), sub_rate=0.08, n_functions=40)
    cases.append(("random", pack(seqs, funcs)))
    seqs, funcs = random_proteins(72, n_families=5, members=(100, 200), length=(100, 400), sub_rate=0.01, alphabet=b"ACDEF")
    cases.append(("heavy duplication", pack(seqs, funcs)))
    cases.append(("tiny", pack(["ACDEFGHIKLMNPQ", "ACDEFGHIKLMNPQ", "ACDEFGHIKLMNPQ", "WWWWWWWWWW"], [0, 0, 0, 1])))
    # fewer proteins than ranks: some ranks encode nothing, some k-mer ranges may be empty
    cases.append(("one protein", pack(["ACDEFGHIKLMNPQRSTVWY"], [3])))
    cases.append(("no valid window", pack(["ACDXFGHIKL", "ACD"], [1, 2])))
    cases.append(("synthetic 40K proteins", Synth(n_proteins=40_000, n_functions=400, n_genomes=8, seed=9).packed()))
    # config-3-shaped function count: 72 000 families, so that the index assignment wraps like the reference's
    # `unsigned short next` (src/function_map.h:324-330), later functions alias earlier indices, proteins whose index
    # would be 0xFFFF are skipped (src/signature_build.tcc:155-158), and both halves of the per-function counters are used
    if not os.environ.get("SIGK_CHECK_SKIP_WRAP"):
        cases.append(("72K functions (index wrap)", Synth(n_proteins=290_000, n_functions=72_000, n_genomes=4, seed=33).packed()))
    for name, p_all in cases:
        print(f"[rank {rank}] case: {name}", file=sys.stderr, flush=True)
        got, per_rank = build_distributed(p_all, rank, world, local_rank)
        if rank == 0:
            from oracle import oracle_c

            want, _ = oracle_c.oracle_build(p_all)
            assert_tables_equal(got, want, tier_b=True, what=name)
            single = GpuSignatureBuilder(device=local_rank)
            single.set_proteins(p_all)
            assert_tables_equal(got, single.build(), tier_b=True, what=name + " vs one GPU")
            single.close()
            print(f"multigpu_check ok ({world} ranks): {name}: {got.n_occurrences} occurrences, kept per rank {per_rank}", flush=True)
        dist.barrier()
    # the encode + route kernel instantiations of the larger worlds (2 and 4 counter words: 5..8 and 9..16 ranks),
    # forced on whatever world size this is
    # (SIGK_SPLIT_KERNEL=warp selects the round-1 per-warp-slice kernel, kept for comparison; the default is the tile-level
    # encode_route_kernel, which has one instantiation for every world size)
    for words in (("1", "2", "4") if os.environ.get("SIGK_CHECK_WIDE") else ()):
        os.environ["SIGK_TEST_SPLIT_WORDS"] = words
        os.environ["SIGK_SPLIT_KERNEL"] = "warp"
        try:
            for name, p_all in (cases[0], cases[5]):
                got, per_rank = build_distributed(p_all, rank, world, local_rank)
                if rank == 0:
                    from oracle import oracle_c

                    want, _ = oracle_c.oracle_build(p_all)
                    assert_tables_equal(got, want, tier_b=True, what=f"{name}, {words} counter words")
                    print(f"multigpu_check ok ({world} ranks): {name} with the per-warp-slice encode+route kernel, {words} counter word(s)", flush=True)
                dist.barrier()
        finally:
            del os.environ["SIGK_TEST_SPLIT_WORDS"]
            del os.environ["SIGK_SPLIT_KERNEL"]
    # regions too small on every rank: the build grows them to the exact need and encodes again
    os.environ["SIGK_TEST_FORCE_SPLIT_FALLBACK"] = "1"
    try:
        name, p_all = cases[0]
        got, per_rank = build_distributed(p_all, rank, world, local_rank)
        if rank == 0:
            from oracle import oracle_c

            want, _ = oracle_c.oracle_build(p_all)
            assert_tables_equal(got, want, tier_b=True, what=f"{name}, forced region overflow")
            print(f"multigpu_check ok ({world} ranks): {name} with a forced region overflow and retry", flush=True)
        dist.barrier()
    finally:
        del os.environ["SIGK_TEST_FORCE_SPLIT_FALLBACK"]
    # without peer mappings: local send regions + NCCL send/recv
    os.environ["SIGK_NO_PEER_WRITES"] = "1"
    try:
        for name, p_all in (cases[0], cases[5]):
            got, per_rank = build_distributed(p_all, rank, world, local_rank)
            if rank == 0:
                from oracle import oracle_c

                want, _ = oracle_c.oracle_build(p_all)
                assert_tables_equal(got, want, tier_b=True, what=f"{name}, NCCL send/recv")
                print(f"multigpu_check ok ({world} ranks): {name} through NCCL send/recv (no peer mappings)", flush=True)
            dist.barrier()
    finally:
        del os.environ["SIGK_NO_PEER_WRITES"]
    # the same inputs, small -> large -> small, through one communicator
    order = [cases[2], cases[5], cases[0]]
    tables = build_sequence_on_one_communicator(order, rank, world, local_rank)
    if rank == 0:
        from oracle import oracle_c

        for (name, p_all), got in zip(order, tables):
            want, _ = oracle_c.oracle_build(p_all)
            assert_tables_equal(got, want, tier_b=True, what=name + " (shared communicator)")
        print(f"multigpu_check ok ({world} ranks): one communicator, growing and shrinking inputs", flush=True)
    dist.barrier()
    if rank == 0:
        print("MULTIGPU_CHECK_PASSED", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
