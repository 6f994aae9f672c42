"""GPU: the reference's JOINT batch semantics (gx_integrate_adaptive_joint; one shared adaptive step for the whole batch,
dynamics/_src/orbit/solver.py:774-803) against the reference's own doctests of that call form, the numpy restatement
(oracle/joint_dopri8.py) and the per-particle kernel."""
import json
from pathlib import Path

import numpy as np
import pytest

import galax_b200.dynamics as gd
import galax_b200.potential as gp
from oracle import joint_dopri8 as jd
from oracle import potentials as op

from conftest import synthetic_ics

pytestmark = pytest.mark.gpu

KATS = json.loads((Path(__file__).parent / "golden" / "orbit_kats.json").read_text())
KMS = KATS["kms"]


def solver(tol=1e-8, **kw):
    return gd.OrbitSolver(solver=kw.pop("method", gd.Dopri8()), stepsize_controller=gd.PIDController(rtol=tol, atol=tol, **kw))


@pytest.mark.parametrize("case", [c for c in KATS["joint"] if c["kind"] == "solve"], ids=lambda c: c["name"][:30])
def test_reference_doctests_of_the_joint_call_form(case):
    """orbit/solver.py:347-380 and :597-613: a batch of two with scalar times is one 12-dimensional ODE; the printed
    8-digit end states are reproduced by the joint kernel (the per-particle kernel differs in the 7th-8th digit)."""
    m, c = case["model"]["params"]
    pot = gp.HernquistPotential(m_tot=m, r_s=c)
    q0, p0 = np.array(case["q0"], float), np.array(case["p0_kms"], float) * KMS
    sol = solver(case["rtol"]).solve(pot, (q0, p0), case["t0"], case["t1"], joint=True)
    q, p = sol.ys[0][:, 0], sol.ys[1][:, 0]
    assert np.allclose(q, case["q"], atol=case["atol"], rtol=0) and np.allclose(p, case["p"], atol=case["atol"], rtol=0)
    n = np.asarray(sol.stats["num_steps"])
    assert n.shape == (2,) and n[0] == n[1] and np.asarray(sol.result).tolist() == [0, 0]


@pytest.mark.parametrize("name,ofun", [("MilkyWayPotential", op.milky_way_potential), ("BovyMWPotential2014", op.bovy_mw_potential_2014)])
def test_joint_kernel_against_the_numpy_restatement(name, ofun):
    pot, opot = getattr(gp, name)(), ofun()
    q0, p0 = synthetic_ics(opot, 150, seed=21)
    ts = np.linspace(0.0, 400.0, 9)
    tol = 1e-9
    qn, pn, stats = jd.solve(opot, q0, p0, 0.0, 400.0, ts, rtol=tol, atol=tol)
    sol = solver(tol).solve(pot, (q0, p0), 0.0, 400.0, saveat=ts, joint=True)
    q, p = sol.ys
    # the shared step sequence is decided by an RMS over 900 numbers: the two codings sum in different orders, so a
    # borderline accept / reject may fall differently -- a handful of steps at most
    assert abs(int(np.asarray(sol.stats["num_steps"])[0]) - stats["num_steps"]) <= 3
    assert abs(int(np.asarray(sol.stats["num_accepted_steps"])[0]) - stats["num_accepted_steps"]) <= 3
    dq = np.abs(q - qn.transpose(1, 0, 2)) / (tol + tol * np.abs(qn.transpose(1, 0, 2)))
    dp = np.abs(p - pn.transpose(1, 0, 2)) / (tol + tol * np.abs(pn.transpose(1, 0, 2)))
    assert dq.max() <= 10.0 and dp.max() <= 10.0, (dq.max(), dp.max())
    # with the step sequence pinned (dt0 given, few steps) the agreement is at rounding level
    qn, pn, stats = jd.solve(opot, q0, p0, 0.0, 20.0, np.array([7.0, 20.0]), rtol=tol, atol=tol, dt0=0.5)
    sol = solver(tol).solve(pot, (q0, p0), 0.0, 20.0, saveat=np.array([7.0, 20.0]), dt0=0.5, joint=True)
    assert int(np.asarray(sol.stats["num_steps"])[0]) == stats["num_steps"]
    assert np.abs(sol.ys[0] - qn.transpose(1, 0, 2)).max() < 1e-11 and np.abs(sol.ys[1] - pn.transpose(1, 0, 2)).max() < 1e-12


def test_joint_mode_properties():
    pot, opot = gp.MilkyWayPotential2022(), op.milky_way_potential_2022()
    q0, p0 = synthetic_ics(opot, 70_001, seed=22)  # more particles than resident threads: every thread walks several
    ts = np.linspace(0.0, 30.0, 4)
    s8 = solver(1e-8)
    a = s8.solve(pot, (q0, p0), 0.0, 30.0, saveat=ts, joint=True)
    b = s8.solve(pot, (q0, p0), 0.0, 30.0, saveat=ts, joint=True)
    assert np.array_equal(a.ys[0], b.ys[0]) and np.array_equal(a.ys[1], b.ys[1])  # fixed-order reductions
    n = np.asarray(a.stats["num_steps"])
    assert (n == n[0]).all() and n[0] > 3
    # per-particle control: the same orbits to the tolerance -- of the JOINT norm, an RMS over 420 006 numbers, which lets
    # the stiffest orbit of the batch err by far more than rtol (measured: 1e-5 kpc worst, 3e-9 typical)
    per = s8.solve(pot, (q0, p0), 0.0, 30.0, saveat=ts)
    dq = np.abs(a.ys[0] - per.ys[0])
    assert dq.max() < 1e-4 and np.median(dq[:, 1:]) < 1e-7 and np.abs(a.ys[1] - per.ys[1]).max() < 1e-5
    assert not np.array_equal(np.asarray(per.stats["num_steps"]), n)
    # one particle: the joint solve IS the per-particle solve (same algorithm, different kernels: rounding only)
    j1 = s8.solve(pot, (q0[:1], p0[:1]), 0.0, 500.0, saveat=np.linspace(0, 500.0, 6), joint=True)
    p1 = s8.solve(pot, (q0[:1], p0[:1]), 0.0, 500.0, saveat=np.linspace(0, 500.0, 6))
    assert abs(int(np.asarray(j1.stats["num_steps"])[0]) - int(np.asarray(p1.stats["num_steps"])[0])) <= 1
    assert np.abs(j1.ys[0] - p1.ys[0]).max() < 1e-7
    # backward in time, the structure-of-arrays layout, Dopri5, a forced minimum step
    qs, ps = q0[:300], p0[:300]
    back = s8.solve(pot, (qs, ps), 0.0, -200.0, saveat=np.linspace(0.0, -200.0, 5), joint=True)
    there = s8.solve(pot, (back.ys[0][:, -1], back.ys[1][:, -1]), -200.0, 0.0, joint=True)
    rt = np.abs(there.ys[0][:, 0] - qs)
    assert rt.max() < 1e-3 and np.median(rt) < 1e-6, (rt.max(), np.median(rt))
    perb = s8.solve(pot, (qs, ps), 0.0, -200.0, saveat=np.linspace(0.0, -200.0, 5))
    assert np.abs(back.ys[0] - perb.ys[0]).max() < 1e-3 and np.median(np.abs(back.ys[0] - perb.ys[0])[:, 1:]) < 1e-6
    t3n = gd._integrate(pot, qs, ps, 0.0, -200.0, np.linspace(0.0, -200.0, 5), solver=gd.Dopri8(), controller=gd.PIDController(1e-8, 1e-8),
                        dt0=None, max_steps=2**16, layout="T3N", joint=True)
    assert np.array_equal(np.asarray(t3n[0]).transpose(2, 0, 1), back.ys[0])
    d5 = solver(1e-7, method=gd.Dopri5()).solve(pot, (qs, ps), 0.0, 100.0, joint=True)
    d8 = s8.solve(pot, (qs, ps), 0.0, 100.0, joint=True)
    d58 = np.abs(d5.ys[0] - d8.ys[0])
    assert d58.max() < 2e-2 and np.median(d58) < 1e-5 and np.asarray(d5.stats["num_steps"])[0] > 20
    forced = solver(1e-12, dtmin=2.0, force_dtmin=True).solve(pot, (qs, ps), 0.0, 100.0, joint=True)
    assert int(np.asarray(forced.stats["num_steps"])[0]) <= 51 and np.isfinite(forced.ys[0]).all()


def test_joint_mode_failures_and_api():
    pot, opot = gp.MilkyWayPotential(), op.milky_way_potential()
    q0, p0 = synthetic_ics(opot, 40, seed=23)
    ts = np.linspace(0.0, 3000.0, 30)
    sol = gd.OrbitSolver(stepsize_controller=gd.PIDController(1e-9, 1e-9), max_steps=25).solve(
        pot, (q0, p0), 0.0, 3000.0, saveat=ts, joint=True, throw=False)
    assert np.asarray(sol.result).tolist() == [1] * 40  # max_steps reached: one status for the one ODE
    nanq = np.isnan(sol.ys[0][..., 0])
    assert nanq.any() and (nanq == nanq[0]).all() and np.isfinite(sol.ys[0][:, 0]).all()
    with pytest.raises(RuntimeError):
        gd.OrbitSolver(max_steps=25).solve(pot, (q0, p0), 0.0, 3000.0, joint=True)
    with pytest.raises(NotImplementedError):  # batched start times are per-particle solves in the reference too
        gd.OrbitSolver().solve(pot, (q0, p0), np.linspace(0, 1, 40), 100.0, joint=True)
    orbit = gd.evaluate_orbit(pot, (q0, p0), np.linspace(0.0, 200.0, 11), joint=True)
    plain = gd.evaluate_orbit(pot, (q0, p0), np.linspace(0.0, 200.0, 11))
    assert orbit.q.shape == (40, 11, 3) and np.abs(orbit.q - plain.q).max() < 1e-3 and not np.array_equal(orbit.q, plain.q)
    co = gd.compute_orbit(pot, (q0, p0), np.linspace(0.0, 200.0, 11), joint=True)
    assert np.abs(co.q - orbit.q).max() < 1e-3
    # a fixed step is shared by construction: joint=True changes nothing
    sie = gd.OrbitSolver(solver=gd.SemiImplicitEuler(), stepsize_controller=gd.ConstantStepSize(), max_steps=None)
    assert np.array_equal(sie.solve(pot, (q0, p0), 0.0, 50.0, dt0=0.5, joint=True).ys[0], sie.solve(pot, (q0, p0), 0.0, 50.0, dt0=0.5).ys[0])
