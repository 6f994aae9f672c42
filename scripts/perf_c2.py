import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent))
import galax_b200.dynamics as gd, galax_b200.potential as gp
from quick_perf import ics, ev_time
pot = gp.MilkyWayPotential2022(); N = 1_000_000; q, p = ics(pot, N, seed=2)
kw = dict(solver=gd.Dopri8(), controller=gd.PIDController(1e-10, 1e-10), dt0=None, max_steps=2**16, throw=False)
for T, layout in ((1, "NT3"), (1000, "NT3"), (1000, "T3N"), (100, "NT3")):
    ts = np.linspace(0, 5000.0, T) if T > 1 else np.array([5000.0])
    out = {}
    def f():
        out["r"] = gd._integrate(pot, q, p, 0.0, 5000.0, ts, layout=layout, **kw)
    best, med = ev_time(f, reps=2, warm=1)
    st = out["r"][3]; na = int(st["num_accepted_steps"].sum())
    print(f"C2 T={T} layout={layout}: {best*1e3:.1f} ms  {na/best:.3e} accepted steps/s")
    del out; torch.cuda.empty_cache()
