"""CPU, world_size 2, gloo: the sharding / gather host logic of the multi-GPU path (no data-path collective
other than the final gather).  The per-shard compute is the CPU oracle here -- only the plumbing is under test."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from galax_b200 import distributed as gdist


def test_shard_bounds_and_dealt_order():
    assert gdist.shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert gdist.shard_bounds(0, 2) == [(0, 0), (0, 0)]
    assert gdist.shard_bounds(3, 8)[:4] == [(0, 1), (1, 2), (2, 3), (3, 3)]
    cost = np.random.default_rng(0).uniform(size=1001)
    perm = gdist.dealt_order(cost, 8)
    assert sorted(perm.tolist()) == list(range(1001))
    loads = [cost[perm[lo:hi]].sum() for lo, hi in gdist.shard_bounds(1001, 8)]
    assert max(loads) / min(loads) < 1.03  # equal cost mix per shard
    plain = [cost[lo:hi].sum() for lo, hi in gdist.shard_bounds(1001, 8)]
    assert max(loads) / min(loads) <= max(plain) / min(plain)


def _worker(rank, world, port, n, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from conftest import synthetic_ics
        from oracle import cref
        from oracle import potentials as op

        opot = op.milky_way_potential()
        q0, p0 = synthetic_ics(opot, n, seed=7)  # same seed on every rank = same global batch
        ts = np.array([0.0, 5.0, 10.0])

        def fn(q, p):
            qq, pp, st, _ = cref.integrate_fixed(opot, q.numpy(), p.numpy(), 0.0, 10.0, 0.1, ts)
            return torch.from_numpy(qq), torch.from_numpy(pp)

        q, p = gdist.integrate_sharded(fn, torch.from_numpy(q0), torch.from_numpy(p0))
        qs, ps = gdist.integrate_sharded(fn, torch.from_numpy(q0), torch.from_numpy(p0), gather=False)
        lo, hi = gdist.shard_bounds(n, world)[rank]
        assert qs.shape[0] == hi - lo and torch.equal(q[lo:hi], qs)
        # the packed, asynchronous form of the same gather (what bench.py uses): two steps in flight
        g = gdist.ResultGather(n, len(ts), device="cpu", depth=2)
        assert g.n_local == hi - lo
        for k in range(3):
            qv, pv = g.views(k)
            qv.copy_(qs + k)
            pv.copy_(ps - k)
            if k >= 2:
                g.wait(k)  # buffer reuse: the gather of step k - 2 must be complete (it was consumed below)
            g.start(k)
            if k >= 1:
                qa, pa = g.result(k - 1)
                assert torch.equal(qa, q + (k - 1)) and torch.equal(pa, p - (k - 1))
        qa, pa = g.result(2)
        assert torch.equal(qa, q + 2) and torch.equal(pa, p - 2)
        torch.save((q, p), f"{tmp}/out{rank}.pt")
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [64, 37, 1])
def test_sharded_integration_gloo_world2(tmp_path, n):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    from conftest import synthetic_ics
    from oracle import cref
    from oracle import potentials as op

    opot = op.milky_way_potential()
    q0, p0 = synthetic_ics(opot, n, seed=7)
    qr, pr, st, _ = cref.integrate_fixed(opot, q0, p0, 0.0, 10.0, 0.1, [0.0, 5.0, 10.0])
    for r in range(2):
        q, p = torch.load(f"{tmp_path}/out{r}.pt")
        assert q.shape == (n, 3, 3)
        assert np.array_equal(q.numpy(), qr) and np.array_equal(p.numpy(), pr)  # sharded == unsharded, bit for bit
