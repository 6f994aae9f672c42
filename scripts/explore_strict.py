"""Strict (reference-order) kernel vs the C oracle, bit for bit, on C1; then the fast kernels against the strict one,
next to the orbit's own sensitivity to a 1-ulp change of its initial condition.  Writes gpurun_out/strict_explore.json."""
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

PAIRS = {
    "MilkyWayPotential": (gp.MilkyWayPotential, op.milky_way_potential),
    "MilkyWayPotential2022": (gp.MilkyWayPotential2022, op.milky_way_potential_2022),
    "BovyMWPotential2014": (gp.BovyMWPotential2014, op.bovy_mw_potential_2014),
}
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
FAST = gd.OrbitSolver(solver=gd.SemiImplicitEuler(), stepsize_controller=gd.ConstantStepSize(), max_steps=None)
STRICT = gd.OrbitSolver(solver=gd.SemiImplicitEuler(strict=True), stepsize_controller=gd.ConstantStepSize(), max_steps=None)


def dev(a, b):
    (qa, pa), (qb, pb) = a, b
    eq = np.linalg.norm(qa - qb, axis=-1) / np.linalg.norm(qb, axis=-1)
    ep = np.linalg.norm(pa - pb, axis=-1) / np.linalg.norm(pb, axis=-1)
    return np.maximum(eq, ep)[:, 0]


out = {}
for name, (cls, ofun) in PAIRS.items():
    pot, opot = cls(), ofun()
    q0, p0 = synthetic_ics(opot, N, seed=1)
    t = time.time()
    s = STRICT.solve(pot, (q0, p0), 0.0, 1000.0, dt0=0.1)
    t_strict = time.time() - t
    t = time.time()
    qr, pr, st, n = cref.integrate_fixed(opot, q0, p0, 0.0, 1000.0, 0.1, [1000.0])
    t_cpu = time.time() - t
    ident = bool(np.array_equal(s.ys[0], qr) and np.array_equal(s.ys[1], pr))
    nbad = int((np.any(s.ys[0] != qr, axis=(1, 2)) | np.any(s.ys[1] != pr, axis=(1, 2))).sum())
    f = FAST.solve(pot, (q0, p0), 0.0, 1000.0, dt0=0.1)
    e = dev(f.ys, s.ys)
    rng = np.random.default_rng(7)
    sens = np.zeros(N)
    for j in range(4):
        sq = rng.integers(0, 2, size=q0.shape) * 2 - 1
        sp = rng.integers(0, 2, size=p0.shape) * 2 - 1
        qj = np.where(sq > 0, np.nextafter(q0, np.inf), np.nextafter(q0, -np.inf))
        pj = np.where(sp > 0, np.nextafter(p0, np.inf), np.nextafter(p0, -np.inf))
        sj = STRICT.solve(pot, (qj, pj), 0.0, 1000.0, dt0=0.1)
        sens = np.maximum(sens, dev(sj.ys, s.ys))
    ratio = e / np.maximum(sens, 1e-17)
    rec = {
        "strict_equals_oracle_bitwise": ident, "particles_differing": nbad, "strict_s": t_strict, "oracle_s": t_cpu,
        "fast_vs_strict": {k: float(np.quantile(e, v)) for k, v in (("median", .5), ("p90", .9), ("p99", .99), ("max", 1))},
        "frac_fast_le_1e-12": float(np.mean(e <= 1e-12)),
        "sens": {k: float(np.quantile(sens, v)) for k, v in (("median", .5), ("p90", .9), ("p99", .99), ("max", 1))},
        "ratio": {k: float(np.quantile(ratio, v)) for k, v in (("median", .5), ("p90", .9), ("p99", .99), ("p999", .999), ("max", 1))},
        "n_fast_gt_1e-12": int((e > 1e-12).sum()),
        "of_those_sens_gt_1e-13": int(((e > 1e-12) & (sens > 1e-13)).sum()),
        "of_those_sens_gt_1e-14": int(((e > 1e-12) & (sens > 1e-14)).sum()),
        "n_sens_gt_1e-13": int((sens > 1e-13).sum()),
        "n_sens_gt_1e-14": int((sens > 1e-14).sum()),
        "worst": [[float(e[i]), float(sens[i])] for i in np.argsort(-e)[:12]],
    }
    out[name] = rec
    print(name, json.dumps(rec), flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/strict_explore.json").write_text(json.dumps(out, indent=1))
