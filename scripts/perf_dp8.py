"""Dopri8 (C2-like) timing for the library selected by $GALAX_B200_LIB."""
import os, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import galax_b200.dynamics as gd, galax_b200.potential as gp
from quick_perf import ev_time, ics
pot = getattr(gp, os.environ.get("POT", "MilkyWayPotential2022"))()
N = int(os.environ.get("N", 148 * 2048)); T = int(os.environ.get("T", 10))
q, p = ics(pot, N, seed=2)
ts = np.linspace(0, 5000.0, T)
kw = dict(solver=gd.Dopri8(), controller=gd.PIDController(1e-10, 1e-10), dt0=None, max_steps=2**16, throw=False)
out = {}
def f():
    out["r"] = gd._integrate(pot, q, p, 0.0, 5000.0, ts, sort=True, **kw)
best, med = ev_time(f, reps=2, warm=1)
st = out["r"][3]
na, nt = st["num_accepted_steps"].sum().item(), st["num_steps"].sum().item()
print(f"{os.environ.get('GALAX_B200_LIB','default')}: N={N} T={T}: {best*1e3:.1f} ms accepted/particle {na/N:.0f} attempted {nt/N:.0f}; {na/best:.3e} acc-steps/s {nt*13/best:.3e} rhs/s")
