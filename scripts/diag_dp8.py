import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import galax_b200.dynamics as gd
import galax_b200.potential as gp
from oracle import cref, potentials as op
from conftest import synthetic_ics
pot, opot = gp.MilkyWayPotential2022(), op.milky_way_potential_2022()
q0, p0 = synthetic_ics(opot, 256, seed=2)
for tol in (1e-10, 1e-7):
  for t1 in (10.0, 100.0, 1000.0, 5000.0):
    solver = gd.OrbitSolver(stepsize_controller=gd.PIDController(rtol=tol, atol=tol), max_steps=2**16)
    ts = np.linspace(0, t1, 5)
    sol = solver.solve(pot, (q0, p0), 0.0, t1, saveat=ts)
    qr, pr, st, na, nt = cref.integrate_dopri8(opot, q0, p0, 0.0, t1, ts, rtol=tol, atol=tol, max_steps=2**16)
    nag = sol.stats["num_accepted_steps"].cpu().numpy(); ntg = sol.stats["num_steps"].cpu().numpy()
    d = np.abs(sol.ys[0]-qr).max(axis=2)   # [N,T]
    print(f"tol={tol} t1={t1}: steps equal frac acc {np.mean(nag==na):.3f} tot {np.mean(ntg==nt):.3f}; |dq| median per save {np.median(d,axis=0)}, max {d.max(axis=0)}")
    same = (nag==na)&(ntg==nt)
    if same.any() and (~same).any():
        print("   same-seq max |dq| final:", d[same,-1].max(), " diff-seq median |dq| final:", np.median(d[~same,-1]))
