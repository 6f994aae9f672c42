"""A/B timing of the save path: fixed step (final only / 101 saves) for the three models, Dopri8 with 10 / 1000 saves.
usage: GALAX_B200_LIB=build_variants/libgx_X.so python scripts/perf_saves.py"""
import json, os, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent))
import galax_b200.dynamics as gd, galax_b200.potential as gp
from quick_perf import ics, ev_time
out = {"lib": os.environ.get("GALAX_B200_LIB", "default")}
SIE = dict(solver=gd.SemiImplicitEuler(), controller=gd.ConstantStepSize(), dt0=0.1, max_steps=None, throw=False)
N = 148 * 8192
for name, cls in (("MW", gp.MilkyWayPotential), ("MW2022", gp.MilkyWayPotential2022), ("Bovy", gp.BovyMWPotential2014)):
    pot = cls(); q, p = ics(pot, N)
    for T in (1, 101):
        ts = np.linspace(0, 1000.0, T) if T > 1 else np.array([1000.0])
        best, med = ev_time(lambda: gd._integrate(pot, q, p, 0.0, 1000.0, ts, **SIE), reps=3)
        out[f"sie_{name}_T{T}_ms"] = round(best * 1e3, 2)
pot = gp.MilkyWayPotential2022(); Nd = 303104; q, p = ics(pot, Nd, seed=2)
kw = dict(solver=gd.Dopri8(), controller=gd.PIDController(1e-10, 1e-10), dt0=None, max_steps=2**16, throw=False)
for T in (1, 10, 1000):
    ts = np.linspace(0, 5000.0, T) if T > 1 else np.array([5000.0])
    best, med = ev_time(lambda: gd._integrate(pot, q, p, 0.0, 5000.0, ts, **kw), reps=3)
    out[f"dp8_MW2022_T{T}_ms"] = round(best * 1e3, 2)
pot = gp.MilkyWayPotential(); q, p = ics(pot, 10_000)
best, med = ev_time(lambda: gd._integrate(pot, q, p, 0.0, 1000.0, np.array([1000.0]), **SIE), reps=5)
out["c1_exact_ms"] = round(best * 1e3, 3)
print(json.dumps(out))
