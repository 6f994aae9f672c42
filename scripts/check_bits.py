"""Prints a digest of the fast kernels' outputs on fixed seeded inputs: run it under two builds of the library
(GALAX_B200_LIB=build_variants/libgx_X.so) to show that a change of instruction sequence did not change a bit."""
import hashlib
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import galax_b200.dynamics as gd  # noqa: E402
import galax_b200.potential as gp  # noqa: E402
from conftest import synthetic_ics  # noqa: E402
from oracle import potentials as op  # noqa: E402  (initial conditions only)


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]


out = {}
for name, ofun in (("MilkyWayPotential", op.milky_way_potential), ("MilkyWayPotential2022", op.milky_way_potential_2022),
                   ("BovyMWPotential2014", op.bovy_mw_potential_2014)):
    pot = getattr(gp, name)()
    q0, p0 = synthetic_ics(ofun(), 20_000, seed=5)
    sie = gd.OrbitSolver(solver=gd.SemiImplicitEuler(), stepsize_controller=gd.ConstantStepSize(), max_steps=None)
    s = sie.solve(pot, (q0, p0), 0.0, 300.0, dt0=0.1, saveat=np.linspace(0.0, 300.0, 7))
    out[name + " fixed"] = digest(*s.ys)
    d8 = gd.OrbitSolver(solver=gd.Dopri8(), stepsize_controller=gd.PIDController(rtol=1e-10, atol=1e-10))
    s = d8.solve(pot, (q0[:5000], p0[:5000]), 0.0, 1000.0, saveat=np.linspace(0.0, 1000.0, 10))
    out[name + " dopri8"] = digest(*s.ys, np.asarray(s.stats["num_steps"]))
comp = gp.CompositePotential(disk=gp.MiyamotoNagaiPotential(m_tot=6.8e10, a=3.0, b=0.28), halo=gp.NFWPotential(m=5.4e11, r_s=15.62),
                             bulge=gp.HernquistPotential(m_tot=5e9, r_s=1.0), extra=gp.PlummerPotential(m_tot=1e9, r_s=0.5))
q0, p0 = synthetic_ics(op.milky_way_potential(), 20_000, seed=6)
s = sie.solve(comp, (q0, p0), 0.0, 300.0, dt0=0.1, saveat=np.linspace(0.0, 300.0, 7))
out["runtime composite fixed"] = digest(*s.ys)
for k, v in out.items():
    print(f"{k:32s} {v}")
