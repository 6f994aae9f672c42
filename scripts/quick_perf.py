"""Quick device timings (CUDA events) of the kernels; development aid, not the bench."""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import galax_b200.dynamics as gd
import galax_b200.potential as gp
from galax_b200 import _lib


def ev_time(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return min(ts), float(np.median(ts))


def ics(pot, N, seed=1):
    g = torch.Generator(device="cuda").manual_seed(seed)
    r = torch.rand(N, generator=g, device="cuda", dtype=torch.float64) * 16 + 4
    d = torch.randn(N, 3, generator=g, device="cuda", dtype=torch.float64)
    d = d / d.norm(dim=1, keepdim=True)
    q = d * r[:, None]
    xr = torch.stack([r, torch.zeros_like(r), torch.zeros_like(r)], 1)
    vc = (r * pot.gradient(xr)[:, 0]).sqrt()
    d2 = torch.randn(N, 3, generator=g, device="cuda", dtype=torch.float64)
    d2 = d2 / d2.norm(dim=1, keepdim=True)
    p = d2 * (vc * (torch.rand(N, generator=g, device="cuda", dtype=torch.float64) * 0.4 + 0.6))[:, None]
    return q.contiguous(), p.contiguous()


def main():
    L = _lib.lib()
    print(torch.cuda.get_device_name(0))
    # DFMA peak
    sink = torch.zeros(8, dtype=torch.float64, device="cuda")
    n = C.c_int64()
    for blocks, threads in ((148 * 8, 256), (148 * 4, 512), (148 * 16, 128)):
        f = lambda: L.gx_bench_dfma(blocks, threads, 20000, sink.data_ptr(), C.byref(n), None)
        best, med = ev_time(f)
        fl = 2.0 * n.value * blocks * threads
        print(f"dfma peak blocks={blocks} threads={threads}: {fl / best / 1e12:.2f} TFLOP/s (median {fl / med / 1e12:.2f})")
    SIE = dict(solver=gd.SemiImplicitEuler(), controller=gd.ConstantStepSize(), dt0=0.1, max_steps=None, throw=False)
    for name, cls in (("MW", gp.MilkyWayPotential), ("MW2022", gp.MilkyWayPotential2022), ("Bovy", gp.BovyMWPotential2014)):
        pot = cls()
        for N in (10_000, 148 * 1024, 148 * 8192):
            q, p = ics(pot, N)
            ts = np.array([1000.0])
            steps = 10000 if N <= 148 * 1024 else 2000
            t1 = steps * 0.1
            f = lambda: gd._integrate(pot, q, p, 0.0, t1, np.array([t1]), **SIE)
            best, med = ev_time(f, reps=2)
            print(f"SIE {name} N={N} steps={steps}: {best * 1e3:.1f} ms  {N * steps / best:.3e} particle-steps/s")
    pot = gp.MilkyWayPotential()
    x = torch.randn(20_000_000, 3, dtype=torch.float64, device="cuda") * 10
    for what, nm, bytes_ in ((_lib.ACC, "acc", 48), (_lib.HESS, "hess", 96), (_lib.ACC | _lib.HESS, "acc+hess", 120), (_lib.PHI, "phi", 32)):
        f = lambda: pot._eval(x, 0.0, what)
        best, med = ev_time(f, reps=3)
        print(f"K1 {nm}: {best * 1e3:.2f} ms  {x.shape[0] / best:.3e} pts/s  {x.shape[0] * bytes_ / best / 1e9:.0f} GB/s")
    # Dopri8 C2-like
    pot = gp.MilkyWayPotential2022()
    for N, T in ((148 * 256, 100), (148 * 1024, 10)):
        q, p = ics(pot, N, seed=2)
        ts = np.linspace(0, 5000.0, T)
        kw = dict(solver=gd.Dopri8(), controller=gd.PIDController(1e-10, 1e-10), dt0=None, max_steps=2**16, throw=False)
        for sort in (True, False):
            out = {}
            def f():
                out["r"] = gd._integrate(pot, q, p, 0.0, 5000.0, ts, sort=sort, **kw)
            best, med = ev_time(f, reps=1, warm=1)
            st = out["r"][3]
            na, nt = st["num_accepted_steps"].sum().item(), st["num_steps"].sum().item()
            print(f"Dopri8 MW2022 N={N} T={T} sort={sort}: {best * 1e3:.1f} ms  accepted {na / N:.0f}/particle attempted {nt / N:.0f}; "
                  f"{na / best:.3e} acc-steps/s {nt * 13 / best:.3e} rhs/s  max/mean steps {st['num_steps'].max().item() / (nt / N):.2f}")


if __name__ == "__main__":
    main()
