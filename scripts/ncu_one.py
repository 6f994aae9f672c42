"""One launch of a bench-shaped kernel for `ncu --set full -k regex:<kernel> -c 1`.
usage: python scripts/ncu_one.py fixed|fixed_bovy|fixed_mw2022|fixed_101|dopri8|dopri8_1000|k1   (env N = particles, LAYOUT = NT3|T3N)"""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent))
import galax_b200.dynamics as gd, galax_b200.potential as gp
from galax_b200 import _lib
from quick_perf import ics
which = sys.argv[1]
import os

N = int(os.environ.get("N", 148 * 8192))
LAYOUT = os.environ.get("LAYOUT", "NT3")
if which == "fixed_101":  # C1's second form: 101 uniform saves
    pot = gp.MilkyWayPotential(); q, p = ics(pot, N, seed=1)
    gd._integrate(pot, q, p, 0.0, 1000.0, np.linspace(0.0, 1000.0, 101), solver=gd.SemiImplicitEuler(),
                  controller=gd.ConstantStepSize(), dt0=0.1, max_steps=None, throw=False, layout=LAYOUT)
elif which == "dopri8_1000":  # C2's shape: 1000 saves over 5 Gyr
    pot = gp.MilkyWayPotential2022(); q, p = ics(pot, N, seed=2)
    gd._integrate(pot, q, p, 0.0, 5000.0, np.linspace(0, 5000.0, 1000), solver=gd.Dopri8(), controller=gd.PIDController(1e-10, 1e-10),
                  dt0=None, max_steps=2**16, throw=False, layout=LAYOUT)
elif which in ("fixed", "fixed_bovy", "fixed_mw2022"):
    pot = {"fixed": gp.MilkyWayPotential, "fixed_bovy": gp.BovyMWPotential2014, "fixed_mw2022": gp.MilkyWayPotential2022}[which]()
    q, p = ics(pot, N, seed=1)
    gd._integrate(pot, q, p, 0.0, 1000.0, np.array([1000.0]), solver=gd.SemiImplicitEuler(), controller=gd.ConstantStepSize(),
                  dt0=0.1, max_steps=None, throw=False)
elif which == "dopri8":
    pot = gp.MilkyWayPotential2022(); q, p = ics(pot, N, seed=2)
    gd._integrate(pot, q, p, 0.0, 5000.0, np.linspace(0, 5000.0, 10), solver=gd.Dopri8(), controller=gd.PIDController(1e-10, 1e-10),
                  dt0=None, max_steps=2**16, throw=False, layout=LAYOUT)
else:
    pot = gp.MilkyWayPotential(); n = 125_000_000
    x = torch.randn(n, 3, dtype=torch.float64, device="cuda") * 10
    pot._eval(x, 0.0, _lib.ACC | _lib.HESS)
torch.cuda.synchronize()
