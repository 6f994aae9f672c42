import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import galax_b200.dynamics as gd
import galax_b200.potential as gp
from oracle import cref, potentials as op
from conftest import synthetic_ics
SIE = gd.OrbitSolver(solver=gd.SemiImplicitEuler(), stepsize_controller=gd.ConstantStepSize(), max_steps=None)
for name,(cls,ofun) in {"MW":(gp.MilkyWayPotential,op.milky_way_potential),"MW2022":(gp.MilkyWayPotential2022,op.milky_way_potential_2022),"Bovy":(gp.BovyMWPotential2014,op.bovy_mw_potential_2014)}.items():
    pot,opot=cls(),ofun()
    q0,p0=synthetic_ics(opot,1024,seed=1)
    sol=SIE.solve(pot,(q0,p0),0.0,1000.0,saveat=np.linspace(0,1000,4001),dt0=0.1)
    rmin=np.linalg.norm(sol.ys[0],axis=2).min(axis=1)
    qr,pr,st,n=cref.integrate_fixed(opot,q0,p0,0.0,1000.0,0.1,[1000.0])
    eq=np.linalg.norm(sol.ys[0][:,-1]-qr[:,0],axis=1)/np.linalg.norm(qr[:,0],axis=1)
    ep=np.linalg.norm(sol.ys[1][:,-1]-pr[:,0],axis=1)/np.linalg.norm(pr[:,0],axis=1)
    e=np.maximum(eq,ep)
    print(name,"median %.2e p90 %.2e p99 %.2e max %.2e"%(np.median(e),np.quantile(e,.9),np.quantile(e,.99),e.max()))
    for cut in (0.1,0.25,0.5,1,2,4):
        m=rmin>cut
        print("   rmin>%.2f: frac %.3f max err %.2e  frac<=1e-12 %.4f"%(cut,m.mean(),e[m].max(),np.mean(e[m]<=1e-12)))
