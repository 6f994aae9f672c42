import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent))
import galax_b200.dynamics as gd, galax_b200.potential as gp
from quick_perf import ics
pot = gp.MilkyWayPotential(); N = 148 * 8192
q, p = ics(pot, N); qh, ph = q.cpu().pin_memory(), p.cpu().pin_memory()
solver = gd.OrbitSolver(solver=gd.SemiImplicitEuler(), stepsize_controller=gd.ConstantStepSize(), max_steps=None)
def step():
    sol = solver.solve(pot, (qh, ph), 0.0, 1000.0, dt0=0.1)
    return sol.ys[0], sol.ys[1]
for i in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    qf, pf = step()
    torch.cuda.synchronize(); print(i, "hold-previous loop ms", (time.perf_counter() - t0) * 1e3)
del qf, pf
for i in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    step()
    torch.cuda.synchronize(); print(i, "discard loop ms", (time.perf_counter() - t0) * 1e3)
for i in range(3):
    t0 = time.perf_counter(); x = torch.empty((N, 1, 3), dtype=torch.float64, pin_memory=True); print("pinned alloc ms", (time.perf_counter() - t0) * 1e3); 
    if i == 1: del x
