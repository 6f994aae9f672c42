"""Round-2 A/B timing of the integrator kernels for the library selected by $GALAX_B200_LIB (CUDA events).
usage: [GALAX_B200_LIB=build_variants/libgx_X.so] python scripts/perf_r2.py [k2] [k3] [k3long]"""
import json, os, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent))
import galax_b200.dynamics as gd, galax_b200.potential as gp
from quick_perf import ev_time, ics
tag = os.path.basename(os.environ.get("GALAX_B200_LIB", "default"))
which = sys.argv[1:] or ["k2", "k3", "k3long"]
res = {"lib": tag}
MODELS = (("MW", gp.MilkyWayPotential), ("MW2022", gp.MilkyWayPotential2022), ("Bovy", gp.BovyMWPotential2014))
if "k2" in which:
    SIE = dict(solver=gd.SemiImplicitEuler(), controller=gd.ConstantStepSize(), dt0=0.1, max_steps=None, throw=False)
    for name, cls in MODELS:
        pot = cls(); N, steps = 148 * 8192, 4000
        q, p = ics(pot, N); t1 = steps * 0.1
        best, med = ev_time(lambda: gd._integrate(pot, q, p, 0.0, t1, np.array([t1]), **SIE), reps=3)
        res[f"k2_{name}_steps_per_s"] = N * steps / best
def dp(name, cls, N, T, tol=1e-10):
    pot = cls(); q, p = ics(pot, N, seed=2); ts = np.linspace(0, 5000.0, T)
    kw = dict(solver=gd.Dopri8(), controller=gd.PIDController(tol, tol), dt0=None, max_steps=2**16, throw=False)
    out = {}
    def f(): out["r"] = gd._integrate(pot, q, p, 0.0, 5000.0, ts, sort=True, **kw)
    best, med = ev_time(f, reps=2, warm=1)
    st = out["r"][3]; nt = st["num_steps"].sum().item()
    res[f"k3_{name}_T{T}_ms"] = best * 1e3; res[f"k3_{name}_T{T}_rhs_per_s"] = nt * 13 / best
if "k3" in which:
    for name, cls in MODELS: dp(name, cls, 148 * 2048, 10)
if "k3long" in which:
    dp("MW2022", gp.MilkyWayPotential2022, 148 * 2048, 1000)
print(json.dumps(res))
