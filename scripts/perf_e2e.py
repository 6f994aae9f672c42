import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent))
import galax_b200.dynamics as gd, galax_b200.potential as gp
from quick_perf import ics
pot = gp.MilkyWayPotential(); N = 148 * 8192
q, p = ics(pot, N); qh, ph = q.cpu().pin_memory(), p.cpu().pin_memory()
kw = dict(solver=gd.SemiImplicitEuler(), controller=gd.ConstantStepSize(), dt0=0.1, max_steps=None)
def T(fn, n=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3
print("device-resident  ms", T(lambda: gd._integrate(pot, q, p, 0.0, 1000.0, [1000.0], **kw)))
print("host (pinned D2H) ms", T(lambda: gd._integrate(pot, qh, ph, 0.0, 1000.0, [1000.0], **kw)))
print("H2D pinned 58MB  ms", T(lambda: (qh.to("cuda"), ph.to("cuda"))))
out = torch.empty((N, 1, 3), dtype=torch.float64, device="cuda")
print("D2H pageable     ms", T(lambda: (out.cpu(), out.cpu())))
pin = torch.empty((N, 1, 3), dtype=torch.float64).pin_memory()
print("D2H pinned       ms", T(lambda: (pin.copy_(out), pin.copy_(out))))
print("empty pageable   ms", T(lambda: (torch.empty((N, 1, 3), dtype=torch.float64).zero_(), torch.empty((N, 1, 3), dtype=torch.float64).zero_())))
