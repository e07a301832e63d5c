"""Customer-sharded runs: one process per GPU (torchrun), contiguous tile-aligned customer ranges,
Philox counters on GLOBAL customer ids, and one all-reduce of the int64 level-2 statistics per sweep
(inside the CUDA library, over NCCL).  torch.distributed is plumbing here: it exchanges the NCCL unique
id and the exact integer partials of the initialisation statistics (works over gloo on CPU as well).
"""
from __future__ import annotations

import numpy as np

from .hostmath import ExactSum

TILE = 1024


def shard_bounds(n_global: int, world: int, tile: int = TILE):
    """[lo, hi) per rank: contiguous, tile-aligned (so the tile -> partial-sum mapping does not depend on
    the number of ranks), remainder tiles spread over the first ranks."""
    ntiles = (n_global + tile - 1) // tile
    base, rem = divmod(ntiles, world)
    out, lo_t = [], 0
    for r in range(world):
        hi_t = lo_t + base + (1 if r < rem else 0)
        out.append((min(lo_t * tile, n_global), min(hi_t * tile, n_global)))
        lo_t = hi_t
    return out


def dist_exact_sum(group=None) -> ExactSum:
    """ExactSum whose reductions run over torch.distributed (any backend)."""
    import torch
    import torch.distributed as dist

    def amax(v: float) -> float:
        objs = [None] * dist.get_world_size(group)
        dist.all_gather_object(objs, float(v), group=group)
        return max(objs)

    def asum(v: int) -> int:
        objs = [None] * dist.get_world_size(group)
        dist.all_gather_object(objs, int(v), group=group)   # Python ints: exact, unbounded
        return sum(objs)

    assert torch is not None
    return ExactSum(amax, asum)


def broadcast_unique_id(make_id, group=None) -> bytes:
    """rank 0 creates the 128-byte NCCL id (make_id()), everyone receives it."""
    import torch.distributed as dist
    obj = [make_id() if dist.get_rank(group) == 0 else None]
    dist.broadcast_object_list(obj, src=0, group=group)
    return obj[0]


def gather_level1(local_draws: np.ndarray, group=None):
    """Concatenate per-rank (n_draws, n_local, ncol) kept draws along customers on rank 0."""
    import torch.distributed as dist
    objs = [None] * dist.get_world_size(group) if dist.get_rank(group) == 0 else None
    dist.gather_object(local_draws, objs, dst=0, group=group)
    if objs is None:
        return None
    return np.concatenate(objs, axis=1)


def connect_p2p(sampler, group=None):
    """Switch a customer-sharded sampler from the NCCL all-reduce to the peer-mailbox all-reduce fused into the
    level-2 kernel: all-gather the CUDA IPC handles of the mailboxes and connect every rank to every other."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if sampler.p2p_is_cached(rank, world):        # every rank made the same calls before: the answer is the same everywhere
        sampler.p2p_connect(None, rank, world)
    else:
        handles = [None] * world
        dist.all_gather_object(handles, sampler.p2p_export(), group=group)
        sampler.p2p_connect(handles, rank, world)
    # every rank has cleared and mapped its mailbox before any rank starts to send (clv_p2p_connect's contract)
    dist.barrier(group=group)
