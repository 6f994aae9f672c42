"""BASELINE.json configurations C3 / C4 / C5 particle-sharded over the GPUs of one node (torchrun), one JSON line.

  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/run_configs_multi.py

Every rank owns a contiguous block (galax_b200.distributed.shard_bounds); the only collective is the all-gather of
result shards (C3, C4); C5's 96 GB of output stays sharded.  Times are the max over ranks of CUDA-synchronised
wall clock around the sharded work + gather.
"""
import json, os, sys, time
from pathlib import Path
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent))
import galax_b200.dynamics as gd, galax_b200.potential as gp
from galax_b200 import _lib, distributed as gdist
from quick_perf import ics

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

def timed(fn):
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return r, float(dt)

out = {"n_gpus": world}
SIE = dict(solver=gd.SemiImplicitEuler(), controller=gd.ConstantStepSize(), dt0=0.1, max_steps=None)

# ---- C4: 1e8 particles, BovyMWPotential2014, SemiImplicitEuler dt = 0.1 Myr, 1 Gyr, final state gathered
pot = gp.BovyMWPotential2014(); N = 100_000_000
lo, hi = gdist.shard_bounds(N, world)[rank]
q, p = ics(pot, hi - lo, seed=4 + rank)
gd._integrate(pot, q[:100000], p[:100000], 0.0, 10.0, np.array([10.0]), **SIE)
def c4():
    qf, pf, st, _ = gd._integrate(pot, q, p, 0.0, 1000.0, np.array([1000.0]), **SIE)
    if world > 1:
        return gdist.all_gather_ragged(qf, N), gdist.all_gather_ragged(pf, N)
    return qf, pf
(qa, pa), dt = timed(c4)
out["C4"] = {"particles": N, "steps": 10_000, "s": dt, "particle_steps_per_s": N * 1e4 / dt, "gathered_shape": list(qa.shape)}
del q, p, qa, pa; torch.cuda.empty_cache()

# ---- C5: 1e9 points, MilkyWayPotential, acceleration + Hessian, outputs stay sharded
pot = gp.MilkyWayPotential(); N = 1_000_000_000
lo, hi = gdist.shard_bounds(N, world)[rank]; n = hi - lo
g = torch.Generator(device="cuda").manual_seed(5 + rank)
r = 10 ** (torch.rand(n, generator=g, device="cuda", dtype=torch.float64) * 3 - 1)
d = torch.randn(n, 3, generator=g, device="cuda", dtype=torch.float64); d /= d.norm(dim=1, keepdim=True)
x = (d * r[:, None]).contiguous(); del d, r
import ctypes as C
P = pot.c_struct(); L = _lib.lib()
acc = torch.empty((n, 3), dtype=torch.float64, device=dev); hess = torch.empty((n, 9), dtype=torch.float64, device=dev)
f5 = lambda: L.gx_potential_eval(C.byref(P), x.data_ptr(), 0.0, n, _lib.ACC | _lib.HESS, None, None, acc.data_ptr(), hess.data_ptr(), None)
f5(); _, dt = timed(f5)
out["C5"] = {"points": N, "s": dt, "points_per_s": N / dt, "GB_per_s_aggregate": N * 120 / dt / 1e9}
del x, acc, hess; torch.cuda.empty_cache()

# ---- C3: Pal-5-like mock stream, Fardal DF, 5e5 stripping times x 2 arms, stripping times sharded
pot = gp.MilkyWayPotential(); M = 500_000
ts = np.linspace(0.0, 3000.0, M)
w0 = gd.PhaseSpaceCoordinate(np.array([30.0, 10, 20]), np.array([10.0, -150, -20]) * gp.KMS, 0.0)
draws = np.random.default_rng(3).standard_normal((4, M))
lo, hi = gdist.shard_bounds(M, world)[rank]
def c3():
    prog = gd.evaluate_orbit(pot, w0, ts)                      # progenitor orbit: redundantly on every rank (12 ms)
    sub = gd.Orbit(torch.as_tensor(prog.q[lo:hi], device=dev), torch.as_tensor(prog.p[lo:hi], device=dev), ts[lo:hi])
    rel = gd.FardalStreamDF().sample(draws[:, lo:hi], pot, sub, 1e4)
    qall = torch.cat([rel["lead"].q, rel["trail"].q]); pall = torch.cat([rel["lead"].p, rel["trail"].p])
    w = gd.Integrator()(gd.HamiltonianField(pot), (qall, pall), np.concatenate([ts[lo:hi], ts[lo:hi]]), float(ts[-1]) + 1e-3)
    m = hi - lo
    if world > 1:
        return [gdist.all_gather_ragged(a.contiguous(), M) for a in (w.q[:m], w.p[:m], w.q[m:], w.p[m:])]
    return [w.q[:m], w.p[:m], w.q[m:], w.p[m:]]
c3(); res, dt = timed(c3)
out["C3"] = {"stripping_times": M, "particles": 2 * M, "s": dt, "released_particles_per_s": 2 * M / dt,
             "finite": bool(torch.isfinite(res[0]).all()), "gathered_shape": list(res[0].shape)}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
