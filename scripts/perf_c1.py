"""C1 exactly (10^4 particles x 10^4 steps, MilkyWayPotential) and the bench-sized batch, for $GALAX_B200_LIB."""
import json, os, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent))
import galax_b200.dynamics as gd, galax_b200.potential as gp
from quick_perf import ev_time, ics
SIE = dict(solver=gd.SemiImplicitEuler(), controller=gd.ConstantStepSize(), dt0=0.1, max_steps=None, throw=False)
res = {"lib": os.path.basename(os.environ.get("GALAX_B200_LIB", "default"))}
for name, cls in (("MW", gp.MilkyWayPotential), ("MW2022", gp.MilkyWayPotential2022), ("Bovy", gp.BovyMWPotential2014)):
    pot = cls()
    q, p = ics(pot, 10_000)
    best, _ = ev_time(lambda: gd._integrate(pot, q, p, 0.0, 1000.0, np.array([1000.0]), **SIE), reps=3)
    res[f"c1_{name}_ms"] = round(best * 1e3, 4)
    q, p = ics(pot, 148 * 8192)
    best, _ = ev_time(lambda: gd._integrate(pot, q, p, 0.0, 400.0, np.array([400.0]), **SIE), reps=3)
    res[f"big_{name}_steps_per_s"] = float(f"{148 * 8192 * 4000 / best:.4g}")
print(json.dumps(res))
