"""Multi-GPU plumbing: one process per GPU, utterances shard embarrassingly.

* exactly one collective, at load: rank 0 reads/folds/packs the voice once and broadcasts the packed
  blob (NCCL over NVLink/NVSwitch on GPUs, gloo on CPU in tests); every rank then builds its own
  device-resident replica from the received bytes.  No collective on the hot path.
* utterance i of a stream goes to rank i % world (round robin) or to the least-loaded rank by
  expected cost (cost ~ number of phonemes, i.e. ~ mel frames); results are re-ordered by sentence
  index, like the reference yields futures in submission order (larynx/__init__.py:187-190).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np


def shard_round_robin(n_items: int, rank: int, world: int) -> List[int]:
    return list(range(rank, n_items, world))


def shard_by_cost(costs: Sequence[float], world: int) -> List[List[int]]:
    """Longest-processing-time-first assignment; deterministic on every rank."""
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    loads = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += float(costs[i])
    return [sorted(x) for x in out]


def bucket_by_length(lengths: Sequence[int], batch: int, max_cost: Optional[int] = None) -> List[List[int]]:
    """Micro-batches of similar length for the fused batch path: sentences sorted by phoneme count
    (stable), cut into runs of at most `batch` items and — if `max_cost` is given — at most `max_cost` phonemes of the
    longest item times the item count (bounds the padded work and the workspace of a micro-batch).  Every index appears
    exactly once; results are put back in submission order by the caller, like the reference yields its futures in
    submission order (larynx/__init__.py:187-190)."""
    order = sorted(range(len(lengths)), key=lambda i: (int(lengths[i]), i))
    out: List[List[int]] = []
    cur: List[int] = []
    for i in order:
        longest = int(lengths[i])  # ascending order: the newcomer is the longest of the run
        if cur and (len(cur) >= batch or (max_cost is not None and longest * (len(cur) + 1) > max_cost)):
            out.append(cur)
            cur = []
        cur.append(i)
    if cur:
        out.append(cur)
    return out


def broadcast_blob(make_blob: Optional[Callable[[], np.ndarray]], device=None, src: int = 0):
    """Rank `src` calls make_blob() (a uint8 numpy array); everyone returns the same bytes as a torch
    uint8 tensor on `device` (cuda:<local_rank> with NCCL, cpu with gloo).  Single-process: no collective."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        t = torch.from_numpy(make_blob())
        return t.to(device) if device is not None else t
    rank = dist.get_rank()
    dev = device if device is not None else "cpu"
    n = torch.zeros(1, dtype=torch.int64, device=dev)
    blob = None
    if rank == src:
        blob = torch.from_numpy(make_blob()).to(dev)
        n[0] = blob.numel()
    dist.broadcast(n, src)
    if rank != src:
        blob = torch.empty(int(n.item()), dtype=torch.uint8, device=dev)
    dist.broadcast(blob, src)
    return blob


def gather_in_order(local_results: dict, n_items: int):
    """All-gather {index: python object} dicts and return a list ordered by utterance index."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        merged = dict(local_results)
    else:
        parts = [None] * dist.get_world_size()
        dist.all_gather_object(parts, local_results)
        merged = {}
        for p in parts:
            merged.update(p)
    assert len(merged) == n_items, (len(merged), n_items)
    return [merged[i] for i in range(n_items)]
