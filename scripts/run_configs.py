"""Run BASELINE.json's five configurations at (single-GPU) full size and print one JSON summary.
C4 / C5 are the per-GPU shards of their 8-GPU statements.  Development / reporting aid (profiles/)."""
import json, sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import galax_b200.dynamics as gd, galax_b200.potential as gp
from galax_b200 import _lib
from quick_perf import ics

def timed(fn, reps=2):
    """Best of ``reps`` after one full-size warm-up (pinned result buffers and the workspace are then cached)."""
    fn(); best = None
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    return r, best

out = {}
which = sys.argv[1:] or ["C1", "C2", "C3", "C4", "C5"]
SIE = dict(solver=gd.SemiImplicitEuler(), controller=gd.ConstantStepSize(), dt0=0.1, max_steps=None)
if "C1" in which:
    pot = gp.MilkyWayPotential(); q, p = ics(pot, 10_000, seed=1)
    gd._integrate(pot, q, p, 0.0, 1000.0, np.array([1000.0]), **SIE)
    r, dt = timed(lambda: gd._integrate(pot, q, p, 0.0, 1000.0, np.array([1000.0]), **SIE))
    r2, dt2 = timed(lambda: gd._integrate(pot, q, p, 0.0, 1000.0, np.linspace(0, 1000, 101), **SIE))
    out["C1"] = {"particles": 10_000, "steps": 10_000, "s_final_only": dt, "particle_steps_per_s": 1e8 / dt, "s_101_saves": dt2}
if "C2" in which:
    pot = gp.MilkyWayPotential2022(); N = 1_000_000; q, p = ics(pot, N, seed=2)
    ts = np.linspace(0, 5000.0, 1000)
    kw = dict(solver=gd.Dopri8(), controller=gd.PIDController(1e-10, 1e-10), dt0=None, max_steps=2**16, throw=False)
    gd._integrate(pot, q[:4096], p[:4096], 0.0, 5000.0, ts, **kw)
    (qq, pp, st, stats), dt = timed(lambda: gd._integrate(pot, q, p, 0.0, 5000.0, ts, **kw))
    na, nt = int(stats["num_accepted_steps"].sum()), int(stats["num_steps"].sum())
    E0 = gd._energy(pot, q, p); E1 = gd._energy(pot, qq[:, -1], pp[:, -1]); drift = (E1 / E0 - 1).abs()
    out["C2"] = {"particles": N, "saves": 1000, "output_GB": 2 * qq.numel() * 8 / 1e9, "s": dt, "accepted_steps": na, "attempted_steps": nt,
                 "accepted_steps_per_s": na / dt, "rhs_per_s": 13 * nt / dt, "failed": int((st != 0).sum()),
                 "energy_drift_median": float(drift.median()), "energy_drift_p99": float(drift.quantile(0.99)) if N <= 16_000_000 else None}
    del qq, pp
    torch.cuda.empty_cache()
if "C3" in which:
    pot = gp.MilkyWayPotential(); M = 500_000
    ts = np.linspace(0.0, 3000.0, M)
    w0 = gd.PhaseSpaceCoordinate(np.array([30.0, 10, 20]), np.array([10.0, -150, -20]) * gp.KMS, 0.0)
    draws = np.random.default_rng(3).standard_normal((4, M))
    gen = gd.MockStreamGenerator(gd.FardalStreamDF(), pot)
    gen.run(draws[:, :1000], ts[:1000], w0, 1e4)
    (stream, prog), dt = timed(lambda: gen.run(draws, ts, w0, 1e4))
    # phase timings
    _, dt_prog = timed(lambda: gd.evaluate_orbit(pot, w0, ts))
    out["C3"] = {"stripping_times": M, "particles": 2 * M, "s_total": dt, "s_progenitor_orbit": dt_prog,
                 "released_particles_per_s": 2 * M / dt, "finite": bool(np.isfinite(stream.q).all())}
if "C4" in which:
    pot = gp.BovyMWPotential2014(); N = 12_500_000; q, p = ics(pot, N, seed=4)
    gd._integrate(pot, q[:100000], p[:100000], 0.0, 10.0, np.array([10.0]), **SIE)
    r, dt = timed(lambda: gd._integrate(pot, q, p, 0.0, 1000.0, np.array([1000.0]), **SIE))
    out["C4_shard"] = {"particles": N, "steps": 10_000, "s": dt, "particle_steps_per_s": N * 1e4 / dt}
if "C5" in which:
    pot = gp.MilkyWayPotential(); N = 125_000_000
    g = torch.Generator(device="cuda").manual_seed(5)
    r = 10 ** (torch.rand(N, generator=g, device="cuda", dtype=torch.float64) * 3 - 1)
    d = torch.randn(N, 3, generator=g, device="cuda", dtype=torch.float64); d /= d.norm(dim=1, keepdim=True)
    x = (d * r[:, None]).contiguous(); del d, r
    pot._eval(x[:1000], 0.0, _lib.ACC | _lib.HESS)
    res, dt = timed(lambda: pot._eval(x, 0.0, _lib.ACC | _lib.HESS))
    out["C5_shard"] = {"points": N, "s": dt, "points_per_s": N / dt, "GB_per_s": N * 120 / dt / 1e9}
print(json.dumps(out))
