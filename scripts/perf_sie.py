import os, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent))
import galax_b200.dynamics as gd, galax_b200.potential as gp
from quick_perf import ev_time, ics
SIE = dict(solver=gd.SemiImplicitEuler(), controller=gd.ConstantStepSize(), dt0=0.1, max_steps=None, throw=False)
for name, cls in (("MW", gp.MilkyWayPotential), ("MW2022", gp.MilkyWayPotential2022), ("Bovy", gp.BovyMWPotential2014)):
    pot = cls()
    for N, steps in ((10_000, 10000), (148 * 8192, 4000)):
        q, p = ics(pot, N)
        t1 = steps * 0.1
        f = lambda: gd._integrate(pot, q, p, 0.0, t1, np.array([t1]), **SIE)
        best, med = ev_time(f, reps=3)
        print(f"{os.path.basename(os.environ.get('GALAX_B200_LIB','default'))} SIE {name} N={N}: {best*1e3:.2f} ms {N*steps/best:.4e} steps/s")
