import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent))
import ctypes as C
import galax_b200.potential as gp
from galax_b200 import _lib
from quick_perf import ev_time
L = _lib.lib()
for cls in (gp.MilkyWayPotential, gp.MilkyWayPotential2022, gp.BovyMWPotential2014):
    pot = cls(); P = pot.c_struct()
    for N in (20_000_000, 125_000_000):
        g = torch.Generator(device="cuda").manual_seed(5)
        r = 10 ** (torch.rand(N, generator=g, device="cuda", dtype=torch.float64) * 3 - 1)
        d = torch.randn(N, 3, generator=g, device="cuda", dtype=torch.float64); d /= d.norm(dim=1, keepdim=True)
        x = (d * r[:, None]).contiguous(); del d, r
        acc = torch.empty((N, 3), dtype=torch.float64, device="cuda"); hess = torch.empty((N, 9), dtype=torch.float64, device="cuda"); phi = torch.empty((N,), dtype=torch.float64, device="cuda")
        for what, nm, b in ((_lib.ACC, "acc", 48), (_lib.HESS, "hess", 96), (_lib.ACC | _lib.HESS, "acc+hess", 120), (_lib.PHI, "phi", 32)):
            f = lambda: L.gx_potential_eval(C.byref(P), x.data_ptr(), 0.0, N, what, phi.data_ptr(), None, acc.data_ptr(), hess.data_ptr(), None)
            best, med = ev_time(f, reps=5, warm=2)
            print(f"{cls.__name__} N={N} {nm}: {best*1e3:.3f} ms {N/best:.3e} pts/s {N*b/best/1e9:.0f} GB/s ({N*b/best/1e9/6531.9*100:.0f}% of measured HBM peak)")
        del x, acc, hess, phi; torch.cuda.empty_cache()
        if cls is not gp.MilkyWayPotential: break
