"""C2-shaped Dopri8 parity: strict (reference-order) kernel vs the C oracle bit for bit; the fast kernel against the strict
one, next to the strict result's own sensitivity to a 1-ulp change of the initial condition.
Writes gpurun_out/strict_dopri_explore.json.   usage: python scripts/explore_strict_dopri.py [N]"""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import galax_b200.dynamics as gd
import galax_b200.potential as gp
from conftest import synthetic_ics
from oracle import cref
from oracle import potentials as op

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
TOL = 1e-10
ts = np.linspace(0.0, 5000.0, 1000)
pot, opot = gp.MilkyWayPotential2022(), op.milky_way_potential_2022()
q0, p0 = synthetic_ics(opot, N, seed=2)
FAST = gd.OrbitSolver(solver=gd.Dopri8(), stepsize_controller=gd.PIDController(TOL, TOL), max_steps=2**16)
STRICT = gd.OrbitSolver(solver=gd.Dopri8(strict=True), stepsize_controller=gd.PIDController(TOL, TOL), max_steps=2**16)


def tolunits(a, ref):
    """max over saves and components of |a - ref| / (atol + rtol |ref|), per particle (q and p)."""
    d = 0.0
    for x, r in zip(a, ref):
        d = np.maximum(d, (np.abs(x - r) / (TOL + TOL * np.abs(r))).max(axis=(1, 2)))
    return d


out = {}
t = time.time()
s = STRICT.solve(pot, (q0, p0), 0.0, 5000.0, saveat=ts)
out["strict_s"] = time.time() - t
t = time.time()
qr, pr, st, na, nt = cref.integrate_dopri8(opot, q0, p0, 0.0, 5000.0, ts, rtol=TOL, atol=TOL, max_steps=2**16)
out["oracle_s"] = time.time() - t
out["strict_equals_oracle_bitwise"] = bool(np.array_equal(s.ys[0], qr) and np.array_equal(s.ys[1], pr))
out["particles_differing"] = int((np.any(s.ys[0] != qr, axis=(1, 2)) | np.any(s.ys[1] != pr, axis=(1, 2))).sum())
out["step_counts_equal"] = bool(np.array_equal(np.asarray(s.stats["num_steps"]), nt) and np.array_equal(np.asarray(s.stats["num_accepted_steps"]), na))
print(json.dumps(out), flush=True)
f = FAST.solve(pot, (q0, p0), 0.0, 5000.0, saveat=ts)
e = tolunits(f.ys, (qr, pr))
same = (np.asarray(f.stats["num_steps"]) == nt) & (np.asarray(f.stats["num_accepted_steps"]) == na)
rng = np.random.default_rng(7)
sens = np.zeros(N)
same_s = np.ones(N, bool)
for j in range(3):
    up = rng.integers(0, 2, size=q0.shape).astype(bool)
    qj = np.where(up, np.nextafter(q0, np.inf), np.nextafter(q0, -np.inf))
    up = rng.integers(0, 2, size=p0.shape).astype(bool)
    pj = np.where(up, np.nextafter(p0, np.inf), np.nextafter(p0, -np.inf))
    sj = STRICT.solve(pot, (qj, pj), 0.0, 5000.0, saveat=ts)
    sens = np.maximum(sens, tolunits(sj.ys, (qr, pr)))
    same_s &= (np.asarray(sj.stats["num_steps"]) == nt) & (np.asarray(sj.stats["num_accepted_steps"]) == na)
qs = lambda x: {k: float(np.quantile(x, v)) for k, v in (("median", .5), ("p90", .9), ("p99", .99), ("max", 1))}  # noqa: E731
out.update({
    "fast_vs_oracle_tolunits": qs(e), "one_ulp_sens_tolunits": qs(sens), "ratio": qs(e / np.maximum(sens, 1e-6)),
    "fast_same_step_counts_frac": float(same.mean()), "one_ulp_same_step_counts_frac": float(same_s.mean()),
    "fast_within_10tol_frac": float((e <= 10).mean()), "one_ulp_within_10tol_frac": float((sens <= 10).mean()),
    "fast_within_10tol_given_same_steps": float((e[same] <= 10).mean()) if same.any() else None,
    "steps_total_fast_over_oracle": float(np.asarray(f.stats["num_steps"]).sum() / nt.sum()),
})
print(json.dumps(out), flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/strict_dopri_explore.json").write_text(json.dumps(out, indent=1))

# how the deviation grows along the orbit: per save-time window, fast vs the one-ulp twin (last draw), tolerance units
def tolunits_t(a, ref):
    d = 0.0
    for x, r in zip(a, ref):
        d = np.maximum(d, (np.abs(x - r) / (TOL + TOL * np.abs(r))).max(axis=2))
    return d  # [N, T]


et, st_ = tolunits_t(f.ys, (qr, pr)), tolunits_t(sj.ys, (qr, pr))
rows = []
for k in (1, 5, 10, 20, 50, 100, 200, 500, 999):
    rows.append({"save": k, "t_myr": float(ts[k]), "fast_median": float(np.median(et[:, k])), "fast_p99": float(np.quantile(et[:, k], .99)),
                 "fast_frac_le_10": float((et[:, : k + 1].max(axis=1) <= 10).mean()), "twin_median": float(np.median(st_[:, k])),
                 "twin_p99": float(np.quantile(st_[:, k], .99)), "twin_frac_le_10": float((st_[:, : k + 1].max(axis=1) <= 10).mean())})
    print(json.dumps(rows[-1]), flush=True)
out["growth"] = rows
Path("gpurun_out/strict_dopri_explore.json").write_text(json.dumps(out, indent=1))
