"""acc + Hessian (C5 shard) and correctness check vs the default library, for $GALAX_B200_LIB."""
import os, sys, ctypes as C
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent))
import galax_b200.potential as gp
from galax_b200 import _lib
from quick_perf import ev_time
L = _lib.lib()
pot = gp.MilkyWayPotential(); P = pot.c_struct()
N = 125_000_000
g = torch.Generator(device="cuda").manual_seed(5)
r = 10 ** (torch.rand(N, generator=g, device="cuda", dtype=torch.float64) * 3 - 1)
d = torch.randn(N, 3, generator=g, device="cuda", dtype=torch.float64); d /= d.norm(dim=1, keepdim=True)
x = (d * r[:, None]).contiguous(); del d, r
acc = torch.empty((N, 3), dtype=torch.float64, device="cuda"); hess = torch.empty((N, 9), dtype=torch.float64, device="cuda")
f = lambda: L.gx_potential_eval(C.byref(P), x.data_ptr(), 0.0, N, _lib.ACC | _lib.HESS, None, None, acc.data_ptr(), hess.data_ptr(), None)
best, med = ev_time(f, reps=7, warm=2)
chk = (float(acc[::9973].sum()), float(hess[::9973].sum()), float(acc[-1, 2]), float(hess[-1, 8]))
print(f"{os.path.basename(os.environ.get('GALAX_B200_LIB','default'))}: acc+hess {best*1e3:.3f} ms {N*120/best/1e9:.0f} GB/s  checksum {chk}")
# ragged size + small size
for n in (1000, 125_000_001 // 1000):
    f2 = lambda: L.gx_potential_eval(C.byref(P), x.data_ptr(), 0.0, n, _lib.ACC | _lib.HESS | _lib.GRAD, None, hess.data_ptr(), acc.data_ptr(), hess[n:].data_ptr(), None)
    f2(); torch.cuda.synchronize()
    print("  n", n, float(acc[:n].sum()), float(hess[n:2 * n].sum()))
