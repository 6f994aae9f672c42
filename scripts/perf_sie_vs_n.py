"""Fixed-step throughput against batch size (MilkyWayPotential, 2000 steps): where the latency-bound regime ends."""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent))
import galax_b200.dynamics as gd, galax_b200.potential as gp
from quick_perf import ev_time, ics
SIE = dict(solver=gd.SemiImplicitEuler(), controller=gd.ConstantStepSize(), dt0=0.1, max_steps=None, throw=False)
pot = gp.MilkyWayPotential()
for N in (1_000, 10_000, 19_000, 30_000, 47_000, 60_000, 75_000, 76_000, 100_000, 150_000, 300_000, 1_000_000, 3_000_000):
    q, p = ics(pot, N)
    f = lambda: gd._integrate(pot, q, p, 0.0, 200.0, np.array([200.0]), **SIE)
    best, med = ev_time(f, reps=3)
    print(f"N={N:8d}: {best*1e3:8.3f} ms  {N*2000/best:.3e} particle-steps/s")
