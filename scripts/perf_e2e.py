"""End-to-end fixed-step rate (pinned host tensors in, host tensors out) against PIPELINE_CHUNKS."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent))
import galax_b200.dynamics as gd, galax_b200.potential as gp
from quick_perf import ics
pot = gp.MilkyWayPotential(); N = 148 * 8192
q, p = ics(pot, N)
qh, ph = q.cpu().pin_memory(), p.cpu().pin_memory()
solver = gd.OrbitSolver(solver=gd.SemiImplicitEuler(), stepsize_controller=gd.ConstantStepSize(), max_steps=None)
for chunks in (4, 6, 8, 12, 16):
    gd.PIPELINE_CHUNKS = chunks
    solver.solve(pot, (qh, ph), 0.0, 1000.0, dt0=0.1)
    best = 1e9
    for _ in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sol = solver.solve(pot, (qh, ph), 0.0, 1000.0, dt0=0.1)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    print(f"chunks={chunks}: {best*1e3:.2f} ms  {N*1e4/best:.4e} particle-steps/s")
