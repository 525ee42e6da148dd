"""Seed-block sharding across GPUs.

Seed blocks are independent (falcon_kit/mains/consensus.py:274 maps them with no shared state), so
the multi-GPU path is a partition with no data-path collective: rank r of N takes a contiguous,
cost-balanced slice of the block list, runs the whole pipeline on its own GPU and the results are
merged in seed order (the ordering contract of ``imap``).  ``torch.distributed`` is used only for
the rendezvous / barrier / timing reduction and, when one rank holds the input, for a broadcast of
the read pool.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def block_costs(block_lens: Sequence[int], seed_lens: Sequence[int]) -> np.ndarray:
    """Cost model: pairs x seed length (DP work grows with both)."""
    return np.asarray(block_lens, dtype=np.float64) * np.asarray(seed_lens, dtype=np.float64)


def partition(costs: Sequence[float], world_size: int) -> List[Tuple[int, int]]:
    """Contiguous [begin, end) slices with near-equal total cost; every block in exactly one slice."""
    n = len(costs)
    if world_size <= 1:
        return [(0, n)]
    c = np.cumsum(np.asarray(costs, dtype=np.float64))
    total = float(c[-1]) if n else 0.0
    cuts = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        k = int(np.searchsorted(c, target, side="left")) + 1 if n else 0
        k = max(k, cuts[-1])
        cuts.append(min(k, n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]


def merge_in_order(parts: Sequence[Tuple[int, Sequence[bytes]]]) -> List[bytes]:
    """parts: (begin index, results) per rank -> results of all blocks in seed order."""
    out: List[bytes] = []
    for begin, res in sorted(parts, key=lambda x: x[0]):
        assert begin == len(out), "shards must tile the block list"
        out.extend(res)
    return out


def gather_results(local: List[bytes], begin: int, world_size: int) -> List[bytes]:
    """All-gather per-rank consensus lists (host objects) and merge them in seed order."""
    if world_size <= 1:
        return list(local)
    import torch.distributed as dist
    objs = [None] * world_size
    dist.all_gather_object(objs, (begin, list(local)))
    return merge_in_order(objs)
