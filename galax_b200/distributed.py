"""Particle sharding across the GPUs of one node (one process per GPU, ``torch.distributed``).

Every particle (orbit, stream particle, evaluation point) is independent, so the path shards with no
data-path collective: each rank integrates a contiguous block of the batch and the only exchange is the
gather of the result shards at the end (NCCL all-gather over NVLink on GPUs; the same code runs on the
``gloo`` backend with CPU tensors, which is how the host logic is tested without GPUs).

The reference has no multi-device path at all (no pmap / shard_map / collectives).
"""

from __future__ import annotations

from typing import Callable, Sequence

import numpy as np


def shard_bounds(n: int, world: int) -> list[tuple[int, int]]:
    """Contiguous blocks of ceil/floor(n/world): the first ``n % world`` ranks get one extra element."""
    if world < 1 or n < 0:
        raise ValueError("world >= 1 and n >= 0 required")
    base, extra = divmod(n, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def dealt_order(cost: np.ndarray, world: int) -> np.ndarray:
    """Permutation that sorts by decreasing cost and deals the particles round-robin over the ranks, so that
    every contiguous shard of the permuted batch gets the same cost mix."""
    order = np.argsort(-np.asarray(cost), kind="stable")
    n = order.shape[0]
    bounds = shard_bounds(n, world)
    out = np.empty(n, dtype=np.int64)
    for r, (lo, hi) in enumerate(bounds):
        out[lo:hi] = order[r::world][: hi - lo]
    # round-robin dealing gives rank r the elements r, r+world, ...; their count matches shard_bounds
    return out


def local_shard(x, rank: int, world: int, dim: int = 0):
    lo, hi = shard_bounds(x.shape[dim], world)[rank]
    idx = [slice(None)] * x.ndim
    idx[dim] = slice(lo, hi)
    return x[tuple(idx)]


def all_gather_ragged(local, n_total: int, group=None, dim: int = 0):
    """All-gather shards whose sizes follow ``shard_bounds(n_total, world)`` along ``dim`` (torch tensors).

    One collective: shards are padded to the largest size, gathered with ``all_gather_into_tensor`` (NCCL:
    a single ring/NVLS all-gather), and the padding is dropped when reassembling.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    bounds = shard_bounds(n_total, world)
    max_len = max(hi - lo for lo, hi in bounds)
    loc = local.movedim(dim, 0).contiguous()
    pad = max_len - loc.shape[0]
    if pad:
        loc = torch.cat([loc, loc.new_zeros((pad, *loc.shape[1:]))], dim=0)
    out = loc.new_empty((world * max_len, *loc.shape[1:]))
    dist.all_gather_into_tensor(out, loc, group=group)
    parts = [out[r * max_len : r * max_len + (hi - lo)] for r, (lo, hi) in enumerate(bounds)]
    return torch.cat(parts, dim=0).movedim(0, dim)


def integrate_sharded(integrate_fn: Callable, q0, p0, *, group=None, gather: bool = True):
    """Run ``integrate_fn(q_shard, p_shard) -> (q, p)`` on this rank's block and gather the results.

    ``q0``/``p0`` are the full ``(N, 3)`` batch (identical on every rank, e.g. broadcast or generated from a
    shared seed); the return value is the full ``(N, T, 3)`` result on every rank (``gather=True``) or this
    rank's shard.
    """
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = q0.shape[0]
    q, p = integrate_fn(local_shard(q0, rank, world), local_shard(p0, rank, world))
    if not gather or world == 1:
        return q, p
    return all_gather_ragged(q, n, group), all_gather_ragged(p, n, group)
