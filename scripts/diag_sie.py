import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import galax_b200.dynamics as gd
import galax_b200.potential as gp
from oracle import cref, potentials as op
from conftest import synthetic_ics
pot, opot = gp.MilkyWayPotential(), op.milky_way_potential()
q0, p0 = synthetic_ics(opot, 384, seed=1)
SIE = gd.OrbitSolver(solver=gd.SemiImplicitEuler(), stepsize_controller=gd.ConstantStepSize(), max_steps=None)
g = pot.gradient(q0); go = op.gradient(opot, q0)
print("grad rel err max", (np.linalg.norm(g-go,axis=1)/np.linalg.norm(go,axis=1)).max())
for t1 in (1.0, 10.0, 100.0, 300.0, 1000.0):
    sol = SIE.solve(pot, (q0, p0), 0.0, t1, dt0=0.1)
    qr, pr, st, n = cref.integrate_fixed(opot, q0, p0, 0.0, t1, 0.1, [t1])
    e = np.linalg.norm(sol.ys[0][:,0]-qr[:,0],axis=1)/np.linalg.norm(qr[:,0],axis=1)
    srt = np.sort(e)
    print(f"t1={t1}: median {np.median(e):.2e} p90 {srt[int(.9*len(e))]:.2e} p99 {srt[int(.99*len(e))]:.2e} max {e.max():.2e} argmax {e.argmax()}")
# pericentre of worst
sol = SIE.solve(pot, (q0, p0), 0.0, 1000.0, saveat=np.linspace(0,1000,2001), dt0=0.1)
rmin = np.linalg.norm(sol.ys[0],axis=2).min(axis=1)
qr, pr, st, n = cref.integrate_fixed(opot, q0, p0, 0.0, 1000.0, 0.1, [1000.0])
e = np.linalg.norm(sol.ys[0][:,-1]-qr[:,0],axis=1)/np.linalg.norm(qr[:,0],axis=1)
idx = np.argsort(e)[::-1][:10]
print("worst particles: err, rmin"); print(np.c_[e[idx], rmin[idx]])
print("corr: particles with rmin>2:", np.max(e[rmin>2]) if (rmin>2).any() else None, " rmin>1:", np.max(e[rmin>1]), "count rmin<1:", (rmin<1).sum())
