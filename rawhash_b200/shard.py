"""Read sharding across the GPUs of one box.

Reads never interact (reference: every read is one independent `map_worker_for` call,
src/rmap.cpp:389), and the index is read-only, so the multi-GPU layout is: one process per GPU,
the flattened index replicated in each GPU's HBM, every mini-batch cut into `world` CONTIGUOUS
read ranges balanced by raw sample count, and the per-rank record arrays concatenated in rank
order — which is input order, what the reference's ordered pipeline step 2 prints
(src/rmap.cpp:744-797).  The only communication is a handful of 64-bit counters per batch
(all-reduce over NCCL/NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Sequence

import numpy as np


def split_by_samples(lens: Sequence[int], world: int) -> np.ndarray:
    """Boundaries b[0..world] of contiguous read ranges with (nearly) equal total sample counts.

    Rank r maps reads [b[r], b[r+1]).  Deterministic, order preserving, every read in exactly one range."""
    lens = np.asarray(lens, dtype=np.int64)
    n = len(lens)
    if world <= 0:
        raise ValueError("world must be positive")
    b = np.zeros(world + 1, dtype=np.int64)
    b[world] = n
    if n == 0:
        return b
    cs = np.cumsum(lens)
    total = int(cs[-1])
    for r in range(1, world):
        target = total * r / world
        i = int(np.searchsorted(cs, target, side="left"))
        # cut after read i if that lands closer to the target than cutting before it
        if i < n and (cs[i] - target) <= (target - (cs[i - 1] if i else 0)):
            i += 1
        b[r] = min(max(i, b[r - 1]), n)
    return b


def merge_records(per_rank: Sequence[np.ndarray], bounds: np.ndarray) -> np.ndarray:
    """Concatenate per-rank record arrays in rank order, turning batch-local read_idx into global."""
    out = []
    for r, recs in enumerate(per_rank):
        recs = recs.copy()
        recs["read_idx"] += np.uint32(bounds[r])
        out.append(recs)
    return np.concatenate(out) if out else np.zeros(0)


def reduce_counters(counters: dict, device=None) -> dict:
    """Sum of integer counters over all ranks and max of `*_ms` timings (slowest rank defines the step).

    Uses the default torch.distributed group (NCCL over NVLink on GPUs, gloo on CPU); identity when
    torch.distributed is not initialised."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(counters)
    keys = sorted(counters)
    sums = torch.tensor([float(counters[k]) if not k.endswith("_ms") else 0.0 for k in keys], dtype=torch.float64, device=device)
    maxs = torch.tensor([float(counters[k]) if k.endswith("_ms") else 0.0 for k in keys], dtype=torch.float64, device=device)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dist.all_reduce(maxs, op=dist.ReduceOp.MAX)
    out = {}
    for i, k in enumerate(keys):
        out[k] = float(maxs[i]) if k.endswith("_ms") else int(round(float(sums[i])))
    return out
