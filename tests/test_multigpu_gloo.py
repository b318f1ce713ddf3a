"""World-size-2 (and 3) runs of the range-partitioned build on CPU over gloo.

The GPU data path of csrc/comm.cu cannot run here, so each rank plays its part
with numpy: encode its chunk of proteins, sample k-mers, agree on splitters
(cut on the case-folded code, so a k-mer in any case pattern has one owner),
split stably by owner, exchange, and reduce its k-mer range in arrival order;
a rank's rows come in the table order of include/sigk.h (k-mers without a
lower-case residue first).  The first sections of the per-rank tables in rank
order followed by their second sections must equal the single-process oracle
bit for bit, including the order-dependent median/var columns — that is the
property the NCCL path relies on (rank r holds canonical chunk r; blocks arrive
in source-rank order; the split is stable)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle_py
from signature_kmers_b200.multigpu import rank_slice
from tests.util import random_proteins

SYM = {c: i for i, c in enumerate(b"ACDEFGHIKLMNPQRSTVWYacdefghiklmnpqrstvwy")}


def encode(seqs, ordinal_base):
    recs = []  # (code, ordinal, offset16)
    for i, s in enumerate(seqs):
        L = len(s)
        for p in range(L - 7):
            w = s[p:p + 8]
            if all(c in SYM for c in w):
                # group code = base-20 code of the case-folded residues << 8 | case mask (residue j = bit j)
                code, mask = 0, 0
                for j, c in enumerate(w):
                    code = code * 20 + SYM[c] % 20
                    mask |= (SYM[c] >= 20) << j
                recs.append(((code << 8) | mask, ordinal_base + i, (L - p) & 0xFFFF))
    return recs


def reduce_records(recs, funcs, lens, sids):
    """process_kmer_set over records already grouped by arrival order (stable sort by code)."""
    # the main run (case mask 0) sorted on the code, then the side run sorted on code and mask; Python's sort is stable
    recs = sorted(recs, key=lambda r: ((r[0] & 0xFF) != 0, r[0]))
    rows, i = [], 0
    sig = set()
    while i < len(recs):
        j = i
        while j < len(recs) and recs[j][0] == recs[i][0]:
            j += 1
        items = recs[i:j][::-1]                        # newest first
        cnt = {}
        for _, o, _ in items:
            cnt[funcs[o]] = cnt.get(funcs[o], 0) + 1
        best = min((f for f in cnt if cnt[f] == max(cnt.values())))
        bc = cnt[best]
        if not (np.float32(bc) < np.float32(len(items)) * np.float32(0.8)):
            acc = oracle_py.BoostAcc()
            for _, o, _ in items:
                if funcs[o] == best:
                    acc.push(lens[o])
                sig.add(sids[o])
            offs = sorted(r[2] for r in items)
            mean, median, var = acc.results()
            rows.append((recs[i][0], offs[len(offs) // 2], best, mean, median, var))
        i = j
    return rows, sig


def worker(rank, world, port, seqs, funcs, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = rank_slice(len(seqs), rank, world)
    recs = encode(seqs[lo:hi], lo)
    # splitters from evenly spaced samples, identical on every rank
    samp = [recs[k * len(recs) // 64][0] >> 8 for k in range(64)] if recs else [20 ** 8] * 64     # case-folded codes
    allsamp = [None] * world
    dist.all_gather_object(allsamp, samp)
    flat = sorted(x for s in allsamp for x in s)
    split = [flat[(k + 1) * 64] for k in range(world - 1)]
    owner = lambda gcode: sum((gcode >> 8) >= s for s in split)
    outgoing = [[r for r in recs if owner(r[0]) == d] for d in range(world)]      # stable split
    incoming = [None] * world
    for d in range(world):                                                         # the all-to-all
        got = [None] * world
        dist.all_gather_object(got, outgoing[d])
        if d == rank:
            incoming = got                                                         # blocks in source-rank order
    mine = [r for block in incoming for r in block]
    lens = [len(s) for s in seqs]
    rows, sig = reduce_records(mine, funcs, lens, list(range(len(seqs))))
    gathered = [None] * world
    dist.all_gather_object(gathered, (rows, sorted(sig), len(recs)))
    if rank == 0:
        out.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,seed", [(2, 5), (3, 6)])
def test_range_partitioned_build_matches_oracle(world, seed):
    seqs, funcs = random_proteins(seed, n_families=14, members=(2, 9), length=(10, 70), sub_rate=0.08, alphabet=b"ACDEFGHIKL")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, world, port, seqs, funcs, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # first sections (case mask 0) in rank order, then second sections in rank order == table order
    rows = [r for part in gathered for r in part[0] if not r[0] & 0xFF] + [r for part in gathered for r in part[0] if r[0] & 0xFF]
    sig = set(x for part in gathered for x in part[1])
    want, stats = oracle_py.build(seqs, funcs)
    alphabet = "ACDEFGHIKLMNPQRSTVWYacdefghiklmnpqrstvwy"

    def decode(gcode):
        code, mask, s = gcode >> 8, gcode & 0xFF, ""
        for j in range(7, -1, -1):
            s = alphabet[code % 20 + (20 if (mask >> j) & 1 else 0)] + s
            code //= 20
        return s.encode()

    got = [(decode(r[0]),) + tuple(r[1:]) for r in rows]
    assert got == want
    assert [r[0] for r in rows] == sorted((r[0] for r in rows), key=lambda g: ((g & 0xFF) != 0, g))
    assert len(sig) == stats["num_seqs_with_a_signature"]
    assert sum(part[2] for part in gathered) == stats["n_occurrences"]
    assert all(len(part[0]) > 0 for part in gathered), "every rank owns part of the k-mer space"


def test_rank_slices_tile_the_input():
    for n in (0, 1, 7, 100, 2_000_001):
        for world in (1, 2, 3, 8):
            cuts = [rank_slice(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in cuts) - min(b - a for a, b in cuts) <= 1
