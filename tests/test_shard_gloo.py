"""CPU, world_size 2 over gloo: the multi-GPU host logic (contiguous sample-balanced read ranges,
order-preserving merge, counter reduction).  The per-read "mapping" here is a stand-in checksum —
the GPU call itself is covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest

from rawhash_b200 import shard


def test_split_is_a_balanced_partition():
    rng = np.random.Generator(np.random.PCG64(3))
    for n, world in [(0, 2), (1, 4), (7, 8), (1000, 2), (1000, 8), (5, 8)]:
        lens = rng.integers(1, 90_000, n)
        b = shard.split_by_samples(lens, world)
        assert b[0] == 0 and b[-1] == n and np.all(np.diff(b) >= 0)
        if n >= 50 * world:
            per = np.array([lens[b[r]:b[r + 1]].sum() for r in range(world)])
            assert per.max() - per.min() <= 2 * lens.max()


def test_library_split_is_a_balanced_partition(built):
    """rh_split_by_samples — what the command line (--gpus N) and the in-context worker ranges use — obeys the same
    contract as shard.split_by_samples: ordered, complete, balanced to within one read either side of every cut."""
    from rawhash_b200 import api
    rng = np.random.Generator(np.random.PCG64(4))
    for n, parts in [(0, 2), (1, 4), (7, 8), (1000, 2), (1000, 8), (5, 8), (3000, 3), (64, 1)]:
        lens = rng.integers(1, 90_000, n).astype(np.uint64)
        b = api.split_by_samples(lens, parts).astype(np.int64)
        assert len(b) == parts + 1 and b[0] == 0 and b[-1] == n and np.all(np.diff(b) >= 0)
        if n:
            cs = np.concatenate([[0], np.cumsum(lens)]).astype(np.float64)
            for r in range(1, parts):  # each cut sits at the first read where the running count reaches r/parts of the total
                target = cs[-1] * r / parts
                assert cs[b[r]] >= target and (b[r] == 0 or cs[b[r] - 1] < target)
    # degenerate shapes: one huge read, zero-length reads, more parts than reads
    assert list(api.split_by_samples([10**12, 1, 1, 1], 4)) == [0, 1, 1, 1, 4]
    assert list(api.split_by_samples([0, 0, 0], 2)) == [0, 1, 3]
    assert list(api.split_by_samples([5], 3)) == [0, 1, 1, 1]


def _worker(rank, world, port, lens, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rawhash_b200 import api
    b = shard.split_by_samples(lens, world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    recs = np.zeros(hi - lo, dtype=api.MAPREC_DTYPE)
    recs["read_idx"] = np.arange(hi - lo)
    recs["sl"] = lens[lo:hi]               # stand-in for the mapping result of each read
    recs["mapped"] = (lens[lo:hi] % 3) != 0
    red = shard.reduce_counters({"n_reads": hi - lo, "n_mapped": int(recs["mapped"].sum()), "step_ms": 10.0 * (rank + 1)})
    q.put((rank, recs, red))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_merge_and_counters():
    import torch.multiprocessing as mp
    world = 2
    rng = np.random.Generator(np.random.PCG64(11))
    lens = rng.integers(100, 80_000, 501).astype(np.int64)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, lens, q)) for r in range(world)]
    for p in ps:
        p.start()
    got = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    b = shard.split_by_samples(lens, world)
    merged = shard.merge_records([g[1] for g in got], b)
    assert np.array_equal(merged["read_idx"], np.arange(len(lens)))     # input order preserved
    assert np.array_equal(merged["sl"], lens)
    for _, _, red in got:                                               # every rank sees the same totals
        assert red["n_reads"] == len(lens)
        assert red["n_mapped"] == int(((lens % 3) != 0).sum())
        assert red["step_ms"] == 20.0                                   # max over ranks
