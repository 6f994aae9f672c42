"""Potential-value / energy kernel throughput (K1 direct path, Phi only) for $GALAX_B200_LIB."""
import os, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent))
import galax_b200.potential as gp
from galax_b200 import _lib
from quick_perf import ev_time
n = 50_000_000
x = torch.randn(n, 3, dtype=torch.float64, device="cuda") * 10
for name, cls in (("MW", gp.MilkyWayPotential), ("MW2022", gp.MilkyWayPotential2022)):
    pot = cls()
    best, med = ev_time(lambda: pot._eval(x, 0.0, _lib.PHI), reps=3)
    print(f"{os.path.basename(os.environ.get('GALAX_B200_LIB', 'default'))} Phi {name}: {best*1e3:.2f} ms {n/best:.3e} points/s {n*32/best/1e9:.0f} GB/s")
