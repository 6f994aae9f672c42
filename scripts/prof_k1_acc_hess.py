"""Only acc+Hessian launches on the C5 shard size (for ncu)."""
import ctypes as C, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import galax_b200.potential as gp
from galax_b200 import _lib
L = _lib.lib(); pot = gp.MilkyWayPotential(); P = pot.c_struct(); N = 125_000_000
g = torch.Generator(device="cuda").manual_seed(5)
r = 10 ** (torch.rand(N, generator=g, device="cuda", dtype=torch.float64) * 3 - 1)
d = torch.randn(N, 3, generator=g, device="cuda", dtype=torch.float64); d /= d.norm(dim=1, keepdim=True)
x = (d * r[:, None]).contiguous(); del d, r
acc = torch.empty((N, 3), dtype=torch.float64, device="cuda"); hess = torch.empty((N, 9), dtype=torch.float64, device="cuda")
for _ in range(4):
    L.gx_potential_eval(C.byref(P), x.data_ptr(), 0.0, N, _lib.ACC | _lib.HESS, None, None, acc.data_ptr(), hess.data_ptr(), None)
torch.cuda.synchronize()
