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
    import torch

    qp = all_gather_ragged(torch.stack([q, p], dim=0), n, group, dim=1)  # ONE collective for both halves
    return qp[0], qp[1]


class ResultGather:
    """The job's only collective, once per step and off the critical path.

    A rank's result shard -- q and p, each ``(n_local, T, 3)`` -- lives in ONE buffer ``[2, n_max, T, 3]`` whose two
    halves the integrator kernels write directly (``views(k)``), so a step needs a single ``all_gather_into_tensor``
    instead of one per half.  ``start(k)`` issues it asynchronously (NCCL: on the communicator's own stream, ordered
    after the kernels already queued on the current stream), so the gather of step k runs under the kernel of step
    k + 1; ``depth`` buffers alternate so that a running gather is never overwritten.  Shards may be ragged
    (``shard_bounds``): every rank's slot is ``n_max`` rows, ``result(k)`` drops the padding.

    Works on any backend (``gloo`` with CPU tensors in the tests)."""

    def __init__(self, n_total: int, T: int, *, device, dtype=None, group=None, depth: int = 2):
        import torch
        import torch.distributed as dist

        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.bounds = shard_bounds(n_total, self.world)
        self.n_local = self.bounds[self.rank][1] - self.bounds[self.rank][0]
        self.n_max = max(hi - lo for lo, hi in self.bounds)
        dtype = dtype or torch.float64
        self.local = [torch.zeros((2, self.n_max, T, 3), dtype=dtype, device=device) for _ in range(depth)]
        self.full = [torch.empty((self.world * 2, self.n_max, T, 3), dtype=dtype, device=device) if self.world > 1
                     else None for _ in range(depth)]  # fmt: skip
        self.work = [None] * depth
        self.depth = depth

    def views(self, k: int):
        """(q, p) views of buffer k for this rank's kernels to write: each ``(n_local, T, 3)``, contiguous."""
        buf = self.local[k % self.depth]
        return buf[0, : self.n_local], buf[1, : self.n_local]

    def start(self, k: int):
        """Issue the gather of step k (asynchronously).  Waits first for the gather that last used this buffer."""
        import torch.distributed as dist

        i = k % self.depth
        if self.world == 1:
            return None
        self.work[i] = dist.all_gather_into_tensor(self.full[i], self.local[i], group=self.group, async_op=True)
        return self.work[i]

    def wait(self, k: int | None = None):
        for i in range(self.depth) if k is None else [k % self.depth]:
            if self.work[i] is not None:
                self.work[i].wait()
                self.work[i] = None

    def result(self, k: int):
        """Full ``(N, T, 3)`` q and p of step k on every rank (waits for its gather)."""
        import torch

        i = k % self.depth
        self.wait(k)
        if self.world == 1:
            return self.views(k)
        full = self.full[i].view(self.world, 2, self.n_max, *self.full[i].shape[2:])
        parts = [full[r, :, : hi - lo] for r, (lo, hi) in enumerate(self.bounds)]
        qp = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
        return qp[0], qp[1]
