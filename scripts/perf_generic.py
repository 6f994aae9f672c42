"""Fixed-step throughput of the runtime-composite kernel against the specialised one, same potential (MilkyWayPotential
rebuilt as a generic CompositePotential), and LM10Potential."""
import os, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent))
import galax_b200.dynamics as gd, galax_b200.potential as gp
from quick_perf import ev_time, ics
SIE = dict(solver=gd.SemiImplicitEuler(), controller=gd.ConstantStepSize(), dt0=0.1, max_steps=None, throw=False)
mw = gp.MilkyWayPotential()
generic = gp.CompositePotential(dict(disk=gp.MiyamotoNagaiPotential(6.8e10, 3.0, 0.28), halo=gp.NFWPotential(5.4e11, 15.62),
                                     bulge=gp.HernquistPotential(5e9, 1.0), nucleus=gp.HernquistPotential(1.71e9, 0.07),
                                     extra=gp.KeplerPotential(1e3)))  # (a fifth component keeps it off the MW kernel; one disk)
for name, pot in (("MW static", mw), ("MW + tiny Plummer, runtime kernel", generic), ("LM10", gp.LM10Potential())):
    N, steps = 148 * 8192, 2000
    q, p = ics(mw, N)
    t1 = steps * 0.1
    f = lambda: gd._integrate(pot, q, p, 0.0, t1, np.array([t1]), **SIE)
    best, med = ev_time(f, reps=2)
    print(f"SIE {name}: {best*1e3:.2f} ms {N*steps/best:.4e} steps/s")
    kw = dict(solver=gd.Dopri8(), controller=gd.PIDController(1e-10, 1e-10), dt0=None, max_steps=2**16, throw=False)
    n2 = 148 * 1024
    out = {}
    g = lambda: out.update(r=gd._integrate(pot, q[:n2], p[:n2], 0.0, 2000.0, np.array([2000.0]), sort=True, **kw))
    best, med = ev_time(g, reps=2)
    nt = out["r"][3]["num_steps"].sum().item()
    print(f"Dopri8 {name}: {best*1e3:.2f} ms {nt*13/best:.3e} rhs/s")
