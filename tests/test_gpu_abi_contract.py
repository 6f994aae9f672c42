"""GPU: what include/galax_b200.h promises about the gx_* entries themselves -- enqueue only (capturable into a CUDA
graph, nothing blocks), safe with several potentials on several streams and host threads, alignment handled, the
degenerate disk plane of a b = 0 (Kuzmin) component finite."""
import ctypes as C
import threading

import numpy as np
import pytest

import galax_b200.dynamics as gd
import galax_b200.potential as gp
from galax_b200 import _lib
from oracle import potentials as op

from conftest import synthetic_ics

pytestmark = pytest.mark.gpu


def _adaptive(L, P, pid, dq, dp, ts, q, p, st, na, nt, ws, stream, t1=300.0):
    n, T = dq.shape[0], ts.shape[0]
    rc = L.gx_integrate_adaptive(_lib.SOLVER_DOPRI8, C.byref(P), C.byref(pid), dq.data_ptr(), dp.data_ptr(), n, None,
                                 0.0, t1, ts.data_ptr(), T, 1 << 16, None, _lib.LAYOUT_NT3, q.data_ptr(), p.data_ptr(),
                                 st.data_ptr(), na.data_ptr(), nt.data_ptr(), ws.data_ptr(), stream)  # fmt: skip
    _lib.check(rc, "gx_integrate_adaptive")


def test_adaptive_entry_is_capturable_with_alternating_potentials():
    """gx_integrate_adaptive with two alternating potentials captured into ONE CUDA graph: no
    synchronisation or allocation may happen inside the entry, and every replay must reproduce the eager results bit
    for bit -- also when other work with a third potential runs on another stream at the same time."""
    import torch

    L = _lib.lib()
    pots = [gp.MilkyWayPotential(), gp.MilkyWayPotential2022()]
    Ps = [p.c_struct() for p in pots]
    pid = gd.PIDController(rtol=1e-8, atol=1e-8).c_struct(None)
    q0, p0 = synthetic_ics(op.milky_way_potential(), 2048, seed=41)
    dq, dp = torch.tensor(q0, device="cuda"), torch.tensor(p0, device="cuda")
    ts = torch.tensor(np.linspace(0.0, 300.0, 7), device="cuda")
    n, T = 2048, 7

    def buffers():
        return (torch.empty((n, T, 3), dtype=torch.float64, device="cuda"), torch.empty((n, T, 3), dtype=torch.float64, device="cuda"),
                torch.empty(n, dtype=torch.int32, device="cuda"), torch.empty(n, dtype=torch.int32, device="cuda"),
                torch.empty(n, dtype=torch.int32, device="cuda"), torch.zeros(32, dtype=torch.int64, device="cuda"))  # fmt: skip

    # eager references (also the warm-up that uploads the force tables: the one allocation an entry may make)
    ref = []
    for P in Ps:
        b = buffers()
        _adaptive(L, P, pid, dq, dp, ts, *b, torch.cuda.current_stream().cuda_stream)
        ref.append(b)
    torch.cuda.synchronize()
    assert all(int((b[2] != 0).sum()) == 0 for b in ref)

    seq = [0, 1, 0, 1, 1, 0]
    outs = [buffers() for _ in seq]
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        s = torch.cuda.current_stream().cuda_stream
        for which, b in zip(seq, outs):
            _adaptive(L, Ps[which], pid, dq, dp, ts, *b, s)
    other = torch.cuda.Stream()
    potC = gp.HernquistPotential(1e12, 5.0)
    solver = gd.OrbitSolver(stepsize_controller=gd.PIDController(rtol=1e-8, atol=1e-8))
    refC = solver.solve(potC, (dq, dp), 0.0, 300.0).ys[0].clone()
    for rep in range(3):
        for b in outs:
            b[0].zero_()
            b[1].zero_()
        torch.cuda.synchronize()
        with torch.cuda.stream(other):  # a third potential in flight on another stream while the graph replays
            oc = solver.solve(potC, (dq, dp), 0.0, 300.0, throw=False).ys[0]
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(oc, refC)
        for which, b in zip(seq, outs):
            assert torch.equal(b[0], ref[which][0]) and torch.equal(b[1], ref[which][1]), (rep, which)
            assert torch.equal(b[3], ref[which][3]) and torch.equal(b[4], ref[which][4])


def test_adaptive_entry_from_several_host_threads():
    """Two host threads, each with its own stream and its own potential, hammering the adaptive entry: every result must
    belong to the caller's potential (the decision "may this launch read the shared constant image" and the launch
    itself are one critical section)."""
    import torch

    pots = [gp.MilkyWayPotential(), gp.BovyMWPotential2014(), gp.HernquistPotential(1e12, 5.0)]
    q0, p0 = synthetic_ics(op.milky_way_potential(), 1024, seed=43)
    dq, dp = torch.tensor(q0, device="cuda"), torch.tensor(p0, device="cuda")
    solver = gd.OrbitSolver(stepsize_controller=gd.PIDController(rtol=1e-8, atol=1e-8))
    refs = [solver.solve(p, (dq, dp), 0.0, 200.0).ys[0].clone() for p in pots]
    torch.cuda.synchronize()
    bad: list = []

    def worker(i):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for rep in range(12):
                k = (i + rep) % len(pots) if rep % 3 == 0 else i
                out = solver.solve(pots[k], (dq, dp), 0.0, 200.0, throw=False).ys[0]
                s.synchronize()
                if not torch.equal(out, refs[k]):
                    bad.append((i, rep, k))

    th = [threading.Thread(target=worker, args=(i,)) for i in range(3)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not bad, bad


def test_bulk_eval_accepts_8_byte_aligned_views():
    """cp.async.bulk (the Hessian path of gx_potential_eval) needs 16-byte aligned addresses; an odd-row view of an
    [N,3] array is only 8-byte aligned and must take the plain path, not fault."""
    import torch

    pot = gp.MilkyWayPotential()
    L, P = _lib.lib(), pot.c_struct()
    x = torch.randn(5001, 3, dtype=torch.float64, device="cuda") * 8

    def acc_hess(xyz, acc, hes):  # the entry itself (the Python mirror would be free to copy a view first)
        rc = L.gx_potential_eval(C.byref(P), xyz.data_ptr(), 0.0, xyz.shape[0], _lib.ACC | _lib.HESS, None, None,
                                 acc.data_ptr(), hes.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "gx_potential_eval")
        torch.cuda.synchronize()

    full_a = torch.empty((5001, 3), dtype=torch.float64, device="cuda")
    full_h = torch.empty((5001, 9), dtype=torch.float64, device="cuda")
    acc_hess(x, full_a, full_h)  # 16-byte aligned: TMA tiles
    view = x[1:]
    assert view.data_ptr() % 16 == 8
    n = view.shape[0]
    acc = torch.empty((n + 1, 3), dtype=torch.float64, device="cuda")[1:]  # misaligned outputs too
    hes = torch.empty((n, 9), dtype=torch.float64, device="cuda")
    acc_hess(view, acc, hes)
    assert torch.equal(acc, full_a[1:]) and torch.equal(hes, full_h[1:])
    assert torch.equal(pot.hessian(view), full_h[1:].reshape(n, 3, 3))
    # a pointer that is not even 8-byte aligned is an argument error, not a fault
    rc = L.gx_potential_eval(C.byref(P), view.data_ptr() + 4, 0.0, 8, _lib.ACC | _lib.HESS, None, None, acc.data_ptr(),
                             hes.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == _lib.lib().gx_version() * 0 - 1  # GX_ERR_BADARG


def test_kuzmin_disk_plane_is_finite():
    """KuzminPotential = Miyamoto-Nagai with b = 0: in the plane z = 0 the closed forms divide by zeta = sqrt(z^2 + b^2)
    = 0.  The reference (builtin/kuzmin.py:82-84, |z|) gives a finite potential, a zero z-force and a finite Hessian
    there; so must the kernels, the oracle, and an orbit that stays in the plane."""
    pot = gp.KuzminPotential(m_tot=1e12, r_s=1.0)
    opot = op.single(0, 1e12, 1.0, 0.0)
    x = np.array([[1.0, 2.0, 0.0], [0.5, 0.0, 0.0], [8.0, -3.0, 0.0], [1.0, 2.0, 3.0], [1.0, 2.0, -1e-200]])
    g, H, phi = pot.gradient(x), pot.hessian(x), pot.potential(x)
    assert np.isfinite(g).all() and np.isfinite(H).all() and np.isfinite(phi).all()
    assert (g[:3, 2] == 0.0).all()
    go, Ho = op.gradient(opot, x), op.hessian(opot, x)
    assert np.abs(g - go).max() <= 1e-14 * np.abs(go).max() and np.abs(H - Ho).max() <= 1e-12 * np.abs(Ho).max()
    # closed form in the plane: grad = GM (R, 0) / (R^2 + a^2)^(3/2)
    R = np.hypot(x[:3, 0], x[:3, 1])
    assert np.allclose(np.hypot(g[:3, 0], g[:3, 1]), pot.G * 1e12 * R / (R**2 + 1.0) ** 1.5, rtol=1e-14)
    # an in-plane orbit stays in the plane and finishes with status OK, fixed step and adaptive; Satoh with b = 0 too
    q0, p0 = np.array([[8.0, 0.0, 0.0]]), np.array([[0.0, 0.2, 0.0]])
    for solver, kw in ((gd.OrbitSolver(solver=gd.SemiImplicitEuler(), stepsize_controller=gd.ConstantStepSize(),
                                       max_steps=None), dict(dt0=0.1)), (gd.OrbitSolver(), {})):  # fmt: skip
        for pt in (pot, gp.SatohPotential(m_tot=1e12, a=1.0, b=0.0)):
            sol = solver.solve(pt, (q0, p0), 0.0, 500.0, **kw)
            assert np.isfinite(sol.ys[0]).all() and (sol.ys[0][..., 2] == 0.0).all() and int(np.asarray(sol.result)[0]) == 0


def test_fardal_chain_entry_is_enqueue_only():
    """gx_jax_fardal_chain takes a caller workspace and can be captured (no malloc / synchronise inside)."""
    import torch

    from galax_b200 import jaxrandom as jr

    L = _lib.lib()
    M = 3000
    draws = torch.empty((4, M), dtype=torch.float64, device="cuda")
    ws = torch.empty(int(L.gx_jax_fardal_chain_workspace_bytes(M)) // 8, dtype=torch.int64, device="cuda")
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        rc = L.gx_jax_fardal_chain(0, 7, M, draws.data_ptr(), ws.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    draws.zero_()
    g.replay()
    torch.cuda.synchronize()
    ref = jr.fardal_draws_per_key(jr.split_chain(jr.key(7), M))
    assert np.allclose(draws.cpu().numpy(), ref, rtol=0, atol=2e-15)
