"""GPU: integrator kernels (K2 fixed step, K3 Dopri8, K4 mock stream) against the CPU oracle, via the C ABI."""
import json
from pathlib import Path

import numpy as np
import pytest

import galax_b200.dynamics as gd
import galax_b200.potential as gp
from oracle import cref
from oracle import potentials as op

from conftest import one_ulp_sensitivity, rel_dev, synthetic_ics

pytestmark = pytest.mark.gpu

KATS = json.loads((Path(__file__).parent / "golden" / "orbit_kats.json").read_text())
KMS = KATS["kms"]
PAIRS = {
    "MilkyWayPotential": (gp.MilkyWayPotential, op.milky_way_potential),
    "MilkyWayPotential2022": (gp.MilkyWayPotential2022, op.milky_way_potential_2022),
    "BovyMWPotential2014": (gp.BovyMWPotential2014, op.bovy_mw_potential_2014),
}
SIE = gd.OrbitSolver(solver=gd.SemiImplicitEuler(), stepsize_controller=gd.ConstantStepSize(), max_steps=None)
LFM = gd.OrbitSolver(solver=gd.LeapfrogMidpoint(), stepsize_controller=gd.ConstantStepSize(), max_steps=None)


def relerr(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def relerr_vec(a, b):
    """max over particles of |a-b| / |b| with |.| the 3-vector norm (component-wise relative error is
    meaningless when a coordinate crosses zero)."""
    return float(np.max(np.linalg.norm(a - b, axis=-1) / np.linalg.norm(b, axis=-1)))


# ------------------------------------------------------------------ fixed step

@pytest.mark.parametrize("name", list(PAIRS))
def test_sie_c1_parity(name):
    """C1 shape: dt = 0.1 Myr over 1 Gyr = 10 000 steps; north_star tolerance 1e-12 relative on final q, p.

    No orbit is excluded by hand.  The bar is 1e-12 for every particle except where the ORACLE's own arithmetic
    (the reference-order kernel, bit-identical to it: tests/test_gpu_strict.py runs all 10^4 C1 particles) moves by more
    than 1e-14 under a one-ulp change of that particle's initial condition; there the bar is 100 x that movement."""
    cls, ofun = PAIRS[name]
    pot, opot = cls(), ofun()
    q0, p0 = synthetic_ics(opot, 384, seed=1)
    sol = SIE.solve(pot, (q0, p0), 0.0, 1000.0, dt0=0.1)
    qr, pr, st, n = cref.integrate_fixed(opot, q0, p0, 0.0, 1000.0, 0.1, [1000.0])
    assert (n == 10000).all()
    sens, strict = one_ulp_sensitivity(pot, q0, p0, 0.0, 1000.0, 0.1)
    assert np.array_equal(strict.ys[0], qr) and np.array_equal(strict.ys[1], pr)
    e = rel_dev(sol.ys, (qr, pr))
    assert (e <= np.maximum(1e-12, 100.0 * sens)).all(), (e / np.maximum(sens, 1e-17)).max()
    assert np.median(e) <= min(1e-13, np.median(sens)) and np.mean(e <= 1e-12) >= np.mean(sens <= 1e-12) - 0.02


def test_sie_saves_layouts_and_interpolation():
    pot, opot = gp.MilkyWayPotential(), op.milky_way_potential()
    q0, p0 = synthetic_ics(opot, 200, seed=2)
    ts = np.concatenate([[0.0], np.sort(np.random.default_rng(0).uniform(0, 300, 37)), [300.0]])
    sol = SIE.solve(pot, (q0, p0), 0.0, 300.0, saveat=ts, dt0=0.25)
    qr, pr, st, n = cref.integrate_fixed(opot, q0, p0, 0.0, 300.0, 0.25, ts)
    assert sol.ys[0].shape == (200, 39, 3)
    assert np.array_equal(sol.ys[0][:, 0], q0) and np.array_equal(sol.ys[1][:, 0], p0)
    assert relerr_vec(sol.ys[0], qr) <= 1e-12 and relerr_vec(sol.ys[1], pr) <= 1e-12
    # structure-of-arrays layout gives the same numbers
    q2, p2, _, _ = gd._integrate(pot, q0, p0, 0.0, 300.0, ts, solver=gd.SemiImplicitEuler(),
                                 controller=gd.ConstantStepSize(), dt0=0.25, max_steps=None, layout="T3N")
    assert q2.shape == (39, 3, 200)
    assert np.array_equal(np.transpose(q2, (2, 0, 1)), sol.ys[0]) and np.array_equal(np.transpose(p2, (2, 0, 1)), sol.ys[1])


def test_sie_kepler_doctest_and_backward_and_ragged():
    case = KATS["cases"][0]
    pot = gp.KeplerPotential(m_tot=1e11)
    sol = SIE.solve(pot, (np.array([case["q0"]]), np.array([case["p0_kms"]]) * KMS), 0.0, 200.0, dt0=0.001,
                    max_steps=200_000)
    assert np.allclose(sol.ys[0][0], case["q"], atol=6e-4) and np.allclose(sol.ys[1][0], case["p"], atol=6e-4)
    mw, omw = gp.MilkyWayPotential(), op.milky_way_potential()
    for N in (1, 31, 33, 1000):  # ragged block sizes
        q0, p0 = synthetic_ics(omw, N, seed=N)
        sol = SIE.solve(mw, (q0, p0), 0.0, -50.0, dt0=-0.1)
        qr, pr, st, n = cref.integrate_fixed(omw, q0, p0, 0.0, -50.0, -0.1, [-50.0])
        assert relerr_vec(sol.ys[0], qr) <= 1e-13 and relerr_vec(sol.ys[1], pr) <= 1e-13
    empty = SIE.solve(mw, (np.zeros((0, 3)), np.zeros((0, 3))), 0.0, 1.0, dt0=0.1)
    assert empty.ys[0].shape == (0, 1, 3)


def test_leapfrog_midpoint_and_max_steps():
    pot, opot = gp.MilkyWayPotential2022(), op.milky_way_potential_2022()
    q0, p0 = synthetic_ics(opot, 100, seed=3)
    sol = LFM.solve(pot, (q0, p0), 0.0, 100.0, dt0=0.05)
    qr, pr, st, n = cref.integrate_fixed(opot, q0, p0, 0.0, 100.0, 0.05, [100.0], scheme=1)
    assert relerr_vec(sol.ys[0], qr) <= 1e-12 and relerr_vec(sol.ys[1], pr) <= 1e-12
    with pytest.raises(RuntimeError, match="max_steps"):
        SIE.solve(pot, (q0, p0), 0.0, 100.0, dt0=0.05, max_steps=10)
    s = SIE.solve(pot, (q0, p0), 0.0, 100.0, dt0=0.05, max_steps=10, throw=False)
    assert (s.result == 1).all() and np.isnan(s.ys[0]).all()


@pytest.mark.parametrize("name", list(PAIRS))
def test_run_length_kernel_is_bit_identical_to_the_general_kernel(name):
    """k_integrate_fixed_seg (run-length time grid, no time arithmetic in the hot loop) against k_integrate_fixed:
    same bits for every grid shape -- saves on and between step boundaries, several saves inside one step, a save
    at t0, backward runs, grids that cross t = 0, a clipped last step, max_steps, both layouts."""
    pot_cls, opot_f = PAIRS[name]
    pot, opot = pot_cls(), opot_f()
    q0, p0 = synthetic_ics(opot, 257, seed=11)
    rng = np.random.default_rng(5)
    kw = dict(solver=gd.SemiImplicitEuler(), controller=gd.ConstantStepSize(), throw=False)
    cases = [
        (0.0, 100.0, 0.1, np.array([100.0]), None),
        (0.0, 100.0, 0.1, np.linspace(0.0, 100.0, 1001), None),                       # every step boundary
        (0.0, 37.3, 0.25, np.sort(rng.uniform(0, 37.3, 500)), None),                  # several saves per step
        (-40.0, 55.5, 0.07, np.concatenate([[-40.0], np.sort(rng.uniform(-40, 55.5, 40)), [55.5]]), None),
        (30.0, -20.0, -0.13, np.sort(rng.uniform(-20, 30, 25))[::-1].copy(), None),   # backward through zero
        (1000.0, 1003.0, 1e-3, np.array([1001.5, 1003.0]), None),                     # large t: coarse ulp of the grid
        (0.0, 10.0, 3.0, np.array([2.9, 3.0, 3.1, 9.99, 10.0]), None),                # clipped last step
        (0.0, 100.0, 0.1, np.array([50.0, 100.0]), 600),                              # max_steps reached
        (5.0, 5.0, 0.1, np.array([5.0]), None),                                       # zero length
    ]
    for t0, t1, dt0, ts, ms in cases:
        for layout in ("NT3", "T3N"):
            a = gd._integrate(pot, q0, p0, t0, t1, ts, dt0=dt0, max_steps=ms, layout=layout, **kw)
            b = gd._integrate(pot, q0, p0, t0, t1, ts, dt0=dt0, max_steps=ms, layout=layout, general_kernel=True, **kw)
            assert np.array_equal(a[0], b[0], equal_nan=True) and np.array_equal(a[1], b[1], equal_nan=True), (t0, t1, dt0)
            assert np.array_equal(a[2], b[2])
        if ms is None and t0 != t1:  # and both agree with the oracle
            qr, pr, _, _ = cref.integrate_fixed(opot, q0, p0, t0, t1, dt0, ts)
            a = gd._integrate(pot, q0, p0, t0, t1, ts, dt0=dt0, max_steps=ms, **kw)
            assert relerr_vec(a[0], qr) <= 1e-11 and relerr_vec(a[1], pr) <= 1e-11


def test_conserved_quantities_full_size():
    """Size-independent properties at a large N: bounded energy error of the symplectic map and exact (to
    rounding) conservation of L_z in the axisymmetric potential."""
    pot, opot = gp.MilkyWayPotential(), op.milky_way_potential()
    q0, p0 = synthetic_ics(opot, 200_000, seed=4)
    sol = SIE.solve(pot, (q0, p0), 0.0, 1000.0, dt0=0.1)
    E0 = gd._energy(pot, q0, p0)
    E1 = gd._energy(pot, sol.ys[0][:, 0], sol.ys[1][:, 0])
    drift = np.abs(E1 / E0 - 1)
    assert np.isfinite(drift).all() and np.median(drift) < 2e-3 and np.quantile(drift, 0.99) < 0.05
    # L_z is conserved exactly by the flow of an axisymmetric potential and to rounding by the map
    Lz0 = q0[:, 0] * p0[:, 1] - q0[:, 1] * p0[:, 0]
    q1, p1 = sol.ys[0][:, 0], sol.ys[1][:, 0]
    Lz1 = q1[:, 0] * p1[:, 1] - q1[:, 1] * p1[:, 0]
    assert np.abs(Lz1 - Lz0).max() < 1e-11 * np.abs(Lz0).max()


# ------------------------------------------------------------------ Dopri8

@pytest.mark.parametrize("case", KATS["cases"][1:], ids=lambda c: c["name"])
def test_reference_dopri8_doctests(case):
    pot = gp.HernquistPotential(*case["model"]["params"])
    solver = gd.OrbitSolver(stepsize_controller=gd.PIDController(rtol=case["rtol"], atol=case["atol_solver"]))
    sol = solver.solve(pot, (np.array(case["q0"]), np.array(case["p0_kms"]) * KMS), case["t0"], case["t1"],
                       saveat=case["ts"])
    assert np.allclose(sol.ys[0], case["q"], atol=case["atol"], rtol=0)
    assert np.allclose(sol.ys[1], case["p"], atol=case["atol"], rtol=0)


def test_reference_single_dopri8_step_doctest():
    """`OrbitSolver.step` (orbit/solver.py:290-321): ONE Dopri8 step of 10 Myr, printed to 8 digits -- pins the
    kernel's tableau arithmetic independently of step control (loose tolerance + dt0 = t1 - t0 => one step)."""
    case = KATS["joint"][0]
    pot = gp.HernquistPotential(*case["model"]["params"])
    solver = gd.OrbitSolver(stepsize_controller=gd.PIDController(rtol=1e-2, atol=1e-2))
    sol = solver.solve(pot, (np.array(case["q0"]), np.array(case["p0_kms"]) * KMS), case["t0"], case["t1"], dt0=10.0)
    assert int(sol.stats["num_steps"].max()) == 1
    assert np.allclose(np.asarray(sol.ys[0])[:, 0], case["q"], atol=case["atol"], rtol=0)
    assert np.allclose(np.asarray(sol.ys[1])[:, 0], case["p"], atol=case["atol"], rtol=0)


@pytest.mark.parametrize("case", KATS["integrate_field"], ids=lambda c: c["name"][:40])
def test_reference_integrate_field_doctest_given_its_first_step(case):
    """dynamics/_src/solver.py:341-365 (Kepler, dtmin = 0.05); see the note in orbit_kats.json."""
    pot = gp.KeplerPotential(*case["model"]["params"])
    solver = gd.OrbitSolver(stepsize_controller=gd.PIDController(rtol=case["rtol"], atol=case["atol_solver"],
                                                                 dtmin=case["dtmin"]), max_steps=case["max_steps"])
    ts = np.linspace(case["t0"], case["t1"], case["n_saves"])
    sol = solver.solve(pot, (np.array(case["q0"]), np.array(case["p0"])), case["t0"], case["t1"], saveat=ts,
                       dt0=case["dt0_observed"])
    for row, ref in case["rows"].items():
        assert np.allclose(sol.ys[0][int(row)], ref["q"], atol=ref["atol"], rtol=0)
        assert np.allclose(sol.ys[1][int(row)], ref["p"], atol=ref["atol"], rtol=0)


@pytest.mark.parametrize("name,tol", [("MilkyWayPotential2022", 1e-10), ("MilkyWayPotential", 1e-7),
                                      ("BovyMWPotential2014", 1e-8)])
def test_dopri8_parity_with_oracle(name, tol):
    """C2 shape (scaled down): saves over 1 Gyr.  north_star bar: |delta| <= 10 (atol + rtol |ref|) at every save.

    diffrax's controller is itself sensitive to rounding: during start-up the error estimate is dominated by
    rounding noise, so two correct implementations (or the oracle on inputs perturbed by 1e-15) pick different
    step sequences for ~10-20% of the particles, and then differ by the method's own global / dense-output
    error (measured: oracle-vs-perturbed-oracle median 4e-11, GPU-vs-oracle median 2e-10 at tol = 1e-10).
    So: (1) the typical particle meets the bar, (2) the GPU's error against a tight-tolerance truth is
    statistically the oracle's error, (3) step counts agree statistically, (4) energy drift is reported.
    """
    cls, ofun = PAIRS[name]
    pot, opot = cls(), ofun()
    q0, p0 = synthetic_ics(opot, 256, seed=2)
    ts = np.linspace(0.0, 1000.0, 21)
    solver = gd.OrbitSolver(stepsize_controller=gd.PIDController(rtol=tol, atol=tol), max_steps=2**16)
    sol = solver.solve(pot, (q0, p0), 0.0, 1000.0, saveat=ts)
    qr, pr, st, na, nt = cref.integrate_dopri8(opot, q0, p0, 0.0, 1000.0, ts, rtol=tol, atol=tol, max_steps=2**16)
    assert (st == 0).all()
    dq = (np.abs(sol.ys[0] - qr) / (tol + tol * np.abs(qr))).max(axis=(1, 2))
    dp = (np.abs(sol.ys[1] - pr) / (tol + tol * np.abs(pr))).max(axis=(1, 2))
    d = np.maximum(dq, dp)
    assert np.median(d) <= 10.0, np.median(d)
    # The reference-order kernel takes the oracle's step sequence exactly (bit-identical saves and step counts); the
    # share of particles the FAST kernel keeps inside the bar is held against what the oracle's own arithmetic keeps
    # inside it when started one ulp away (tests/test_gpu_strict.py has the full-size C2 version and the reasoning).
    strict = gd.OrbitSolver(solver=gd.Dopri8(strict=True), stepsize_controller=gd.PIDController(rtol=tol, atol=tol),
                            max_steps=2**16)  # fmt: skip
    s0 = strict.solve(pot, (q0, p0), 0.0, 1000.0, saveat=ts)
    assert np.array_equal(s0.ys[0], qr) and np.array_equal(s0.ys[1], pr)
    assert np.array_equal(np.asarray(s0.stats["num_steps"]), nt) and np.array_equal(np.asarray(s0.stats["num_accepted_steps"]), na)
    rng = np.random.default_rng(5)
    nudge = lambda x: np.where(rng.integers(0, 2, size=x.shape) > 0, np.nextafter(x, np.inf), np.nextafter(x, -np.inf))  # noqa: E731
    s1 = strict.solve(pot, (nudge(q0), nudge(p0)), 0.0, 1000.0, saveat=ts)
    dt_ = np.maximum((np.abs(s1.ys[0] - qr) / (tol + tol * np.abs(qr))).max(axis=(1, 2)),
                     (np.abs(s1.ys[1] - pr) / (tol + tol * np.abs(pr))).max(axis=(1, 2)))
    assert np.mean(d <= 10.0) >= np.mean(dt_ <= 10.0) - 0.3 and np.median(d) <= 4.0 * np.median(dt_) + 0.5
    qt, pt, stt, _, _ = cref.integrate_dopri8(opot, q0, p0, 0.0, 1000.0, ts, rtol=1e-13, atol=1e-13)
    err_gpu = np.abs(sol.ys[0] - qt).max(axis=(1, 2))
    err_orc = np.abs(qr - qt).max(axis=(1, 2))
    assert 0.5 <= np.median(err_gpu) / np.median(err_orc) <= 2.0
    assert np.quantile(err_gpu, 0.9) <= 3.0 * np.quantile(err_orc, 0.9)
    na_gpu = sol.stats["num_accepted_steps"].cpu().numpy()
    nt_gpu = sol.stats["num_steps"].cpu().numpy()
    assert abs(na_gpu.sum() / na.sum() - 1) < 0.01 and abs(nt_gpu.sum() / nt.sum() - 1) < 0.01
    assert np.mean(na_gpu == na) > 0.5
    E0 = gd._energy(pot, q0, p0)
    E1 = gd._energy(pot, sol.ys[0][:, -1], sol.ys[1][:, -1])
    assert np.quantile(np.abs(E1 / E0 - 1), 0.99) < 1e3 * tol


def test_dopri8_short_horizon_parity():
    """With a given dt0 and a short horizon the GPU and the oracle take (nearly) the same steps."""
    pot, opot = gp.MilkyWayPotential2022(), op.milky_way_potential_2022()
    q0, p0 = synthetic_ics(opot, 256, seed=12)
    ts = np.linspace(0.0, 60.0, 13)
    tol = 1e-10
    solver = gd.OrbitSolver(stepsize_controller=gd.PIDController(rtol=tol, atol=tol), max_steps=2**16)
    sol = solver.solve(pot, (q0, p0), 0.0, 60.0, saveat=ts, dt0=2.0)
    qr, pr, st, na, nt = cref.integrate_dopri8(opot, q0, p0, 0.0, 60.0, ts, rtol=tol, atol=tol, dt0=2.0)
    same = (sol.stats["num_steps"].cpu().numpy() == nt) & (sol.stats["num_accepted_steps"].cpu().numpy() == na)
    assert same.mean() > 0.9
    d = np.maximum((np.abs(sol.ys[0] - qr) / (tol + tol * np.abs(qr))).max(axis=(1, 2)),
                   (np.abs(sol.ys[1] - pr) / (tol + tol * np.abs(pr))).max(axis=(1, 2)))
    # Even with equal step counts the step sizes differ in their last digits (the controller's factor inherits
    # the relative rounding noise of a far-below-tolerance error estimate), so saved values differ by a
    # fraction of the dense-output error; the bar holds for the bulk, not for every particle.
    assert np.median(d) <= 1.0, np.median(d)
    assert np.mean(d <= 10.0) >= 0.8, np.mean(d <= 10.0)
    # step end points (theta = 1, no interpolation) agree far better than the bar
    de = np.abs(sol.ys[0][:, -1] - qr[:, -1]) / (tol + tol * np.abs(qr[:, -1]))
    assert np.median(de.max(axis=1)) <= 0.1


def test_dopri8_step_sequence_is_bitwise_stable_and_order_independent():
    pot, opot = gp.MilkyWayPotential2022(), op.milky_way_potential_2022()
    q0, p0 = synthetic_ics(opot, 3000, seed=8)
    ts = np.linspace(0.0, 1000.0, 11)
    kw = dict(solver=gd.Dopri8(), controller=gd.PIDController(1e-9, 1e-9), dt0=None, max_steps=None)
    a = gd._integrate(pot, q0, p0, 0.0, 1000.0, ts, sort=True, **kw)
    b = gd._integrate(pot, q0, p0, 0.0, 1000.0, ts, sort=False, **kw)
    c = gd._integrate(pot, q0, p0, 0.0, 1000.0, ts, sort=True, layout="T3N", **kw)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.array_equal(np.transpose(c[0], (2, 0, 1)), a[0])


def test_dopri8_per_particle_t0_backward_and_zero_length():
    pot, opot = gp.MilkyWayPotential(), op.milky_way_potential()
    q0, p0 = synthetic_ics(opot, 300, seed=9)
    t0 = np.random.default_rng(1).uniform(0, 2999.0, 300)
    t0[:3] = 3000.0  # zero-length integrations return y0
    integ = gd.Integrator()
    w = integ(gd.HamiltonianField(pot), (q0, p0), t0, 3000.0)
    qr, pr, st, na, nt = cref.integrate_dopri8(opot, q0, p0, t0, 3000.0, [3000.0], rtol=1e-7, atol=1e-7)
    assert np.array_equal(w.q[:3], q0[:3])
    assert np.abs(w.q - qr[:, 0]).max() < 2e-4 and np.abs(w.p - pr[:, 0]).max() < 2e-5
    wb = integ(gd.HamiltonianField(pot), (q0, p0), 0.0, -500.0)
    qb, pb, *_ = cref.integrate_dopri8(opot, q0, p0, 0.0, -500.0, [-500.0], rtol=1e-7, atol=1e-7)
    assert np.abs(wb.q - qb[:, 0]).max() < 2e-4


def test_evaluate_orbit_and_compute_orbit_api():
    pot, opot = gp.MilkyWayPotential(), op.milky_way_potential()
    q0, p0 = synthetic_ics(opot, 64, seed=10)
    t = np.linspace(0.0, 500.0, 26)
    orb = gd.evaluate_orbit(pot, np.concatenate([q0, p0], axis=1), t)
    assert orb.q.shape == (64, 26, 3) and orb.p.shape == (64, 26, 3) and orb.t.shape == (26,)
    assert np.array_equal(orb.q[:, 0], q0)
    qr, pr, *_ = cref.integrate_dopri8(opot, q0, p0, 0.0, 500.0, t, rtol=1e-7, atol=1e-7)
    assert np.abs(orb.q - qr).max() < 1e-4
    # w0 carrying its own time: first integrate w0.t -> t[0] (legacy/funcs.py:194-206)
    w0 = gd.PhaseSpaceCoordinate(q0, p0, -100.0)
    orb2 = pot.evaluate_orbit(w0, t)
    qa, pa, *_ = cref.integrate_dopri8(opot, q0, p0, -100.0, 0.0, [0.0], rtol=1e-7, atol=1e-7)
    assert np.abs(orb2.q[:, 0] - qa[:, 0]).max() < 1e-5
    orb3 = gd.compute_orbit(pot, (q0, p0), t)
    qr8, *_ = cref.integrate_dopri8(opot, q0, p0, 0.0, 500.0, t, rtol=1e-8, atol=1e-8)
    assert np.abs(orb3.q - qr8).max() < 2e-5
    single = pot.evaluate_orbit((q0[0], p0[0]), t)
    assert single.q.shape == (26, 3)
    E = orb3.total_energy()
    assert E.shape == (64, 26) and np.abs(E / E[:, :1] - 1).max() < 1e-6
    fixed = gd.Integrator(dynamics_solver=gd.OrbitSolver(solver=gd.SemiImplicitEuler(),
                                                         stepsize_controller=gd.ConstantStepSize()),
                          diffeq_kw={"max_steps": None, "dt0": 0.1})
    orb4 = gd.evaluate_orbit(pot, (q0, p0), t, integrator=fixed)
    qf, *_ = cref.integrate_fixed(opot, q0, p0, 0.0, 500.0, 0.1, t)
    assert relerr_vec(orb4.q[:, 1:], qf[:, 1:]) < 1e-12


# ------------------------------------------------------------------ mock stream

def test_adaptive_launches_with_different_potentials_on_two_streams():
    """The Dopri8 RHS reads its parameters from one __constant__ image per device (stage_const_pot): interleaved
    launches with different potentials on different streams must each see their own parameters."""
    import torch

    potA, potB = gp.MilkyWayPotential(), gp.HernquistPotential(1e12, 5.0)
    q0, p0 = synthetic_ics(op.milky_way_potential(), 4096, seed=31)
    dq, dp = torch.tensor(q0, device="cuda"), torch.tensor(p0, device="cuda")
    solver = gd.OrbitSolver(stepsize_controller=gd.PIDController(rtol=1e-8, atol=1e-8))
    refA = solver.solve(potA, (dq, dp), 0.0, 300.0).ys[0].clone()
    refB = solver.solve(potB, (dq, dp), 0.0, 300.0).ys[0].clone()
    torch.cuda.synchronize()
    sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for rep in range(4):
        with torch.cuda.stream(sA):
            outs.append(("A", solver.solve(potA, (dq, dp), 0.0, 300.0, throw=False).ys[0]))
        with torch.cuda.stream(sB):
            outs.append(("B", solver.solve(potB, (dq, dp), 0.0, 300.0, throw=False).ys[0]))
    torch.cuda.synchronize()
    for tag, o in outs:
        assert torch.equal(o, refA if tag == "A" else refB), tag


def test_time_dependent_linear_parameters():
    """LinearParameter (potential/_src/params/core.py:25-110): the reference's own doctest (Kepler losing mass, rho at 10
    saves to 3 decimals) through compute_orbit; frozen-time bulk evaluation; fixed-step and Dopri8 orbits of a
    composite with growing disk and halo against the oracle; unsupported combinations fail loudly."""
    case = KATS["time_dependent"][0]
    lp = gp.LinearParameter(slope=case["slope_msun_per_myr"], point_time=case["point_time"], point_value=case["point_value"])
    pot = gp.KeplerPotential(m_tot=lp)
    assert pot.is_time_dependent and not gp.KeplerPotential(1e12).is_time_dependent
    w0 = gd.PhaseSpaceCoordinate(np.array(case["q0"]), np.array(case["p0_kms"]) * KMS, case["t0"])
    orbit = gd.compute_orbit(pot, w0, np.linspace(case["t0"], case["t1"], case["n_saves"]))
    assert np.allclose(np.hypot(orbit.q[:, 0], orbit.q[:, 1]), case["rho"], atol=case["atol"], rtol=0)

    mix = gp.CompositePotential(
        disk=gp.MiyamotoNagaiPotential(gp.LinearParameter(1e7, 0.0, 6.8e10), gp.LinearParameter(1e-4, 100.0, 3.01), 0.28),
        halo=gp.NFWPotential(gp.LinearParameter(2e8, 0.0, 5.4e11), gp.LinearParameter(1e-3, 0.0, 15.62)),
        bulge=gp.HernquistPotential(5e9, 1.0), nuc=gp.JaffePotential(gp.LinearParameter(-1e5, 0.0, 1e9), 0.3),
        iso=gp.IsochronePotential(3e9, gp.LinearParameter(1e-4, 0.0, 2.0)), sat=gp.SatohPotential(2e9, 3.0, gp.LinearParameter(1e-5, 0.0, 0.4)),
        tri=gp.TriaxialHernquistPotential(gp.LinearParameter(1e6, 0.0, 4e9), 0.8, 0.9, gp.LinearParameter(1e-5, 0.0, 0.7)))
    omix = op.Potential((
        op.Component(op.KIND_MN, (6.8e10, 3.0, 0.28), rates=(1e7, 1e-4, 0.0)), op.Component(op.KIND_NFW, (5.4e11, 15.62), rates=(2e8, 1e-3)),
        op.Component(op.KIND_HERNQUIST, (5e9, 1.0)), op.Component(op.KIND_JAFFE, (1e9, 0.3), rates=(-1e5, 0.0)),
        op.Component(op.KIND_ISOCHRONE, (3e9, 2.0), rates=(0.0, 1e-4)), op.Component(op.KIND_SATOH, (2e9, 3.0, 0.4), rates=(0.0, 0.0, 1e-5)),
        op.Component(op.KIND_TRIAXIAL_HERNQUIST, (4e9, 0.8, 0.9, 0.7), rates=(1e6, 0.0, 0.0, 1e-5))))
    xyz = np.random.default_rng(4).normal(size=(2000, 3)) * 8
    for t in (0.0, 730.0):
        go = op.gradient(omix, xyz, t)
        assert (np.abs(mix.gradient(xyz, t) - go) / np.linalg.norm(go, axis=1, keepdims=True)).max() < 5e-15
        Ho = op.hessian(omix, xyz, t)
        assert (np.abs(mix.hessian(xyz, t) - Ho) / np.abs(Ho).max(axis=(1, 2), keepdims=True)).max() < 5e-13
    q0, p0 = synthetic_ics(op.milky_way_potential(), 256, seed=41)
    sie = gd.OrbitSolver(solver=gd.SemiImplicitEuler(), stepsize_controller=gd.ConstantStepSize(), max_steps=None)
    sol = sie.solve(mix, (q0, p0), 0.0, 500.0, dt0=0.1)
    qr, pr, st, n = cref.integrate_fixed(omix, q0, p0, 0.0, 500.0, 0.1, [500.0])
    rel = np.linalg.norm(sol.ys[0][:, 0] - qr[:, 0], axis=-1) / np.linalg.norm(qr[:, 0], axis=-1)
    assert np.median(rel) < 1e-13 and np.mean(rel < 1e-11) > 0.95
    frozen = sie.solve(gp.CompositePotential(disk=gp.MiyamotoNagaiPotential(6.8e10, 3.0, 0.28), halo=gp.NFWPotential(5.4e11, 15.62)),
                       (q0, p0), 0.0, 500.0, dt0=0.1)
    assert np.abs(frozen.ys[0] - sol.ys[0]).max() > 1e-2  # the time dependence matters
    ts = np.linspace(0.0, 500.0, 6)
    for t0, t1, tsv in ((0.0, 500.0, ts), (500.0, 0.0, ts[::-1].copy())):  # forward and backward in time
        so = gd.OrbitSolver(stepsize_controller=gd.PIDController(rtol=1e-9, atol=1e-9)).solve(mix, (q0, p0), t0, t1, saveat=tsv)
        qd, pd, *_ = cref.integrate_dopri8(omix, q0, p0, t0, t1, tsv, rtol=1e-9, atol=1e-9)
        d = np.abs(so.ys[0] - qd).max(axis=(1, 2))
        assert np.median(d) < 1e-7 and np.quantile(d, 0.9) < 1e-5, (t0, np.median(d), np.quantile(d, 0.9))
    # not supported: time-dependent kinds outside the integrators' set, stream release, energies
    bad = gp.CompositePotential(b=gp.PowerLawCutoffPotential(gp.LinearParameter(1e5, 0.0, 5e9), 1.8, 1.9))
    assert np.isfinite(bad.gradient(xyz[:4], 100.0)).all()  # frozen-time evaluation works for every kind
    with pytest.raises(Exception):
        sie.solve(bad, (q0[:4], p0[:4]), 0.0, 10.0, dt0=0.1)
    with pytest.raises(NotImplementedError):
        gp.MN3Sech2Potential(gp.LinearParameter(1.0, 0.0, 4.7e10), 2.6, 0.3)._flat_components()


def test_reference_tidal_radius_doctests_through_release_kernel():
    """cluster/api.py:54-64,70-76,180-198 (lagrange_points / tidal_radius doctests): with all Fardal draws zero the
    release kernel places the leading particle at x - 2 r_t r_hat, so r_t is read off K4 directly."""
    from test_oracle_integrators import TIDAL_KATS

    for name, x, v, mass, rt in TIDAL_KATS:
        pot = gp.MilkyWayPotential() if name == "MilkyWayPotential" else gp.NFWPotential(1e12, 20.0)
        orbit = gd.Orbit(np.array([x]), np.array([v]), np.array([0.0]))
        out = gd.FardalStreamDF().sample(np.zeros((4, 1)), pot, orbit, mass)
        assert abs((x[0] - out["lead"].q[0, 0]) / 2 - rt) < 6e-9 and abs((out["trail"].q[0, 0] - x[0]) / 2 - rt) < 6e-9, name


def test_phase_space_diagnostics_on_device():
    """kinetic / potential / total energy and angular momentum of PhaseSpaceCoordinate and Orbit
    (coordinates/_src/pscs/base.py:182-330; doctest :304-317: q = [1,0,0], p = [0,2,0] -> L = [0,0,2])."""
    w = gd.PhaseSpaceCoordinate(np.array([1.0, 0, 0]), np.array([0, 2.0, 0]), 0.0)
    assert np.array_equal(w.angular_momentum(), [0.0, 0.0, 2.0]) and w.kinetic_energy() == 2.0
    pot, opot = gp.MilkyWayPotential(), op.milky_way_potential()
    q0, p0 = synthetic_ics(opot, 64, seed=21)
    orb = gd.evaluate_orbit(pot, (q0, p0), np.linspace(0.0, 200.0, 9))
    assert orb.q.shape == (64, 9, 3)
    K, U, E, L = orb.kinetic_energy(), orb.potential_energy(), orb.total_energy(), orb.angular_momentum()
    assert K.shape == (64, 9) and L.shape == (64, 9, 3)
    assert np.allclose(K, 0.5 * (orb.p**2).sum(-1), rtol=1e-15) and np.allclose(U, op.potential(opot, orb.q), rtol=1e-13)
    assert np.allclose(E, K + U, rtol=1e-14) and np.allclose(L, np.cross(orb.q, orb.p), rtol=1e-14, atol=1e-16)
    assert np.abs(E / E[:, :1] - 1).max() < 1e-5 and np.abs(L[..., 2] - L[:, :1, 2]).max() < 1e-6  # axisymmetric: E, L_z
    assert np.allclose(gd.PhaseSpacePosition(q0, p0).total_energy(pot), E[:, 0], rtol=1e-14)


@pytest.mark.parametrize("dfname", ["fardal", "chen"])
def test_stream_release_matches_oracle(dfname):
    pot, opot = gp.MilkyWayPotential(), op.milky_way_potential()
    rng = np.random.default_rng(3)
    M = 500
    q0, p0 = synthetic_ics(opot, M, seed=11)
    orbit = gd.Orbit(q0, p0, np.linspace(0, 1, M))
    if dfname == "fardal":
        draws = rng.standard_normal((4, M))
        out = gd.FardalStreamDF().sample(draws, pot, orbit, 1e4)
        ref = cref.release_fardal(opot, q0, p0, 1e4, draws)
    else:
        draws = gd.ChenStreamDF()._draws(5, M).cpu().numpy()  # made on the device (gx_jax_normal)
        out = gd.ChenStreamDF().sample(draws, pot, orbit, 1e4)
        ref = cref.release_chen(opot, q0, p0, 1e4, draws)
    got = (out["lead"].q, out["lead"].p, out["trail"].q, out["trail"].p)
    for g, r in zip(got, ref):
        assert relerr_vec(g, r) < 1e-12


def test_mockstream_generator_matches_oracle_and_reference_test_shape():
    """The reference's own generator test (NFW host, 10 stripping times; shapes + finiteness), then values
    against the oracle on a Milky-Way stream."""
    host = gp.NFWPotential(m=1.0e12, r_s=15.0)
    ts = np.linspace(0.0, 4000.0, 10)
    w0 = gd.PhaseSpaceCoordinate(np.array([30.0, 10, 20]), np.array([10.0, -150, -20]) * KMS, 0.0)
    gen = gd.MockStreamGenerator(gd.FardalStreamDF(), host)
    stream, prog = gen.run(12, ts, w0, 1e4)
    assert stream.q.shape == (20, 3) and stream.p.shape == (20, 3) and np.isfinite(stream.q).all()
    assert stream["lead"].q.shape == (10, 3) and prog.q.shape == (3,)

    pot, opot = gp.MilkyWayPotential(), op.milky_way_potential()
    M = 400
    ts = np.linspace(0.0, 3000.0, M)
    draws = np.random.default_rng(3).standard_normal((4, M))
    gen = gd.MockStreamGenerator(gd.FardalStreamDF(), pot)
    stream, prog = gen.run(draws, ts, w0, 1e4)
    ref = cref.mockstream(opot, [w0.q], [w0.p], ts, 1e4, draws)
    assert np.abs(prog.q - ref["prog_q"][-1]).max() < 2e-4  # two 1e-7 solves with independent step sequences
    # stream particles feel a 1e-7 solve twice (progenitor + particle): agreement at the 1e-4 kpc level
    assert np.abs(stream["lead"].q - ref["lead_q"]).max() < 5e-3
    assert np.abs(stream["trail"].q - ref["trail_q"]).max() < 5e-3
    assert np.median(np.abs(stream["lead"].q - ref["lead_q"])) < 2e-4


def test_mockstream_generator_seeded_draws_chen_and_device_inputs():
    """Integer seeds follow jax's key chain (Fardal: split + normal; Chen: multivariate_normal(method="svd")); the same
    draws passed explicitly give the same stream; CUDA-tensor progenitors keep the result on the device."""
    import torch

    from galax_b200 import jaxrandom

    # the device generator against the numpy restatement (threefry bits exact, erfinv to a few ulp)
    for seed, n in ((0, 7), (5, 100_003), (2**40 + 17, 4096)):
        k = jaxrandom.key(seed)
        z = gd._device_jax_normal(k, n).cpu().numpy()
        zr = jaxrandom.normal(k, (n,))
        assert np.abs(z - zr).max() < 2e-15 * max(1.0, np.abs(zr).max()) * 4 and abs(z.mean()) < 5 / np.sqrt(n) + 1e-12
    pot = gp.MilkyWayPotential()
    M = 300
    ts = np.linspace(0.0, 2000.0, M)
    w0 = gd.PhaseSpaceCoordinate(np.array([30.0, 10, 20]), np.array([10.0, -150, -20]) * KMS, 0.0)
    for df, draws in ((gd.FardalStreamDF(), jaxrandom.fardal_draws(5, M)), (gd.ChenStreamDF(), jaxrandom.chen_draws(5, M))):
        gen = gd.MockStreamGenerator(df, pot)
        s1, p1 = gen.run(5, ts, w0, 1e4)
        s2, p2 = gen.run(draws, ts, w0, 1e4)
        # seeded draws are made on the device (gx_jax_normal); the host restatement uses scipy's erfinv: equal to a few ulp.
        # A few-ulp difference in a release condition can flip one accept/reject decision of the 1e-7 solve, which moves
        # that particle by a fraction of the tolerance (measured: 2 of 600 particles by 5e-8 kpc, all others < 1e-9).
        dq, dp = np.abs(s1.q - s2.q).max(axis=-1), np.abs(s1.p - s2.p).max(axis=-1)
        assert np.median(dq) < 1e-10 and np.mean(dq < 1e-8) > 0.97 and dq.max() < 1e-6 and dp.max() < 1e-7
        assert np.isfinite(s1.q).all()
        s3, _ = gen.run(6, ts, w0, 1e4)
        assert not np.array_equal(s1.q, s3.q)
        # lead and trail sit on opposite sides of the progenitor's final position, a few tidal radii away
        d_lead = np.linalg.norm(s1["lead"].q[-1] - p1.q), np.linalg.norm(s1["trail"].q[-1] - p1.q)
        assert 0.0 < min(d_lead) and max(d_lead) < 2.0
    wd = gd.PhaseSpaceCoordinate(torch.tensor(w0.q, device="cuda"), torch.tensor(w0.p, device="cuda"), 0.0)
    sd, pd = gd.MockStreamGenerator(gd.FardalStreamDF(), pot).run(5, ts, wd, 1e4)
    assert sd["lead"].q.is_cuda and pd.q.is_cuda
    s1, _ = gd.MockStreamGenerator(gd.FardalStreamDF(), pot).run(5, ts, w0, 1e4)
    assert np.array_equal(sd["lead"].q.cpu().numpy(), s1["lead"].q)


def test_single_orbit_record_and_parallel_dense_output_matches_in_kernel_saves():
    """N = 1 with many saves takes gx_integrate_dopri8_record + gx_dense_eval; it must reproduce the in-kernel
    SaveAt path (same steps, same continuous extension)."""
    pot = gp.MilkyWayPotential()
    q0 = np.array([[30.0, 10.0, 20.0]])
    p0 = np.array([[10.0, -150.0, -20.0]]) * KMS
    ts = np.linspace(0.0, 3000.0, 777)
    kw = dict(solver=gd.Dopri8(), controller=gd.PIDController(1e-7, 1e-7), dt0=None, max_steps=None)
    q1, p1, st1, s1 = gd._integrate(pot, q0, p0, 0.0, 3000.0, ts, **kw)  # N = 1, T >= 64: record path
    q2, p2, st2, s2 = gd._integrate(pot, np.repeat(q0, 2, 0), np.repeat(p0, 2, 0), 0.0, 3000.0, ts, **kw)
    assert q1.shape == (1, 777, 3)
    assert int(s1["num_accepted_steps"][0]) == int(s2["num_accepted_steps"][0])
    assert np.allclose(q1[0], q2[0], rtol=1e-13, atol=1e-13) and np.allclose(p1[0], p2[0], rtol=1e-13, atol=1e-15)
    assert np.array_equal(q1[0, 0], q0[0])
    # backward in time and a record buffer that is too small
    tb = np.linspace(0.0, -1000.0, 100)
    qb1, _, _, _ = gd._integrate(pot, q0, p0, 0.0, -1000.0, tb, **kw)
    qb2, _, _, _ = gd._integrate(pot, np.repeat(q0, 2, 0), np.repeat(p0, 2, 0), 0.0, -1000.0, tb, **kw)
    assert np.allclose(qb1[0], qb2[0], rtol=1e-13, atol=1e-13)
    with pytest.raises(RuntimeError, match="max_steps"):
        gd._integrate(pot, q0, p0, 0.0, 3000.0, ts, solver=gd.Dopri8(), controller=gd.PIDController(1e-7, 1e-7),
                      dt0=None, max_steps=20)


def test_host_arrays_round_trip_through_pinned_memory():
    """numpy in -> numpy out, CPU torch in -> CPU torch out (pinned staging), identical to the CUDA-tensor path."""
    import torch

    pot, opot = gp.MilkyWayPotential(), op.milky_way_potential()
    q0, p0 = synthetic_ics(opot, 5003, seed=21)
    kw = dict(solver=gd.SemiImplicitEuler(), controller=gd.ConstantStepSize(), dt0=0.1, max_steps=None)
    ts = np.array([0.0, 12.5, 50.0])
    ref = gd._integrate(pot, torch.from_numpy(q0).cuda(), torch.from_numpy(p0).cuda(), 0.0, 50.0, ts, **kw)
    assert ref[0].is_cuda
    got = gd._integrate(pot, q0, p0, 0.0, 50.0, ts, **kw)
    assert isinstance(got[0], np.ndarray) and np.array_equal(got[0], ref[0].cpu().numpy())
    got_t = gd._integrate(pot, torch.from_numpy(q0).pin_memory(), torch.from_numpy(p0).pin_memory(), 0.0, 50.0, ts, **kw)
    assert not got_t[0].is_cuda and np.array_equal(got_t[1].numpy(), ref[1].cpu().numpy())
    keep = got[0].copy()
    gd._integrate(pot, q0 * 1.01, p0, 0.0, 50.0, ts, **kw)  # a later call must not overwrite an earlier result
    assert np.array_equal(got[0], keep)


def test_pipelined_host_path_equals_single_launch():
    """Pinned host tensors above PIPELINE_MIN_PARTICLES go through the chunked copy/compute pipeline: same bits as
    one launch on device-resident inputs, for the fixed-step and the adaptive kernels, statuses included."""
    import torch

    pot, opot = gp.MilkyWayPotential(), op.milky_way_potential()
    N = gd.PIPELINE_MIN_PARTICLES + 1237  # ragged chunk bounds
    q0, p0 = synthetic_ics(opot, 4096, seed=21)
    reps = -(-N // 4096)
    scale = 1.0 + 1e-3 * np.arange(reps)[:, None, None]
    q0 = (q0[None] * scale).reshape(-1, 3)[:N].copy()
    p0 = np.tile(p0, (reps, 1))[:N].copy()
    qp, pp = torch.from_numpy(q0).pin_memory(), torch.from_numpy(p0).pin_memory()
    qd, pd = qp.cuda(), pp.cuda()
    ts = np.array([10.0, 20.0])
    kw = dict(solver=gd.SemiImplicitEuler(), controller=gd.ConstantStepSize(), dt0=0.1, max_steps=None)
    a = gd._integrate(pot, qp, pp, 0.0, 20.0, ts, **kw)
    b = gd._integrate(pot, qd, pd, 0.0, 20.0, ts, **kw)
    assert not a[0].is_cuda and a[0].shape == (N, 2, 3)
    assert torch.equal(a[0], b[0].cpu()) and torch.equal(a[1], b[1].cpu()) and torch.equal(a[2], b[2].cpu())
    c = gd._integrate(pot, q0, p0, 0.0, 20.0, ts, **kw)  # numpy (pageable) in -> numpy out, same pipeline
    assert isinstance(c[0], np.ndarray) and np.array_equal(c[0], b[0].cpu().numpy()) and np.array_equal(c[1], b[1].cpu().numpy())
    kw = dict(solver=gd.Dopri8(), controller=gd.PIDController(1e-8, 1e-8), dt0=None, max_steps=4096)
    t0 = np.linspace(0.0, 5.0, N)  # per-particle start times are sliced with the particles
    a = gd._integrate(pot, qp, pp, t0, 20.0, ts, **kw)
    b = gd._integrate(pot, qd, pd, torch.from_numpy(t0).cuda(), 20.0, ts, **kw)
    assert torch.equal(a[0], b[0].cpu()) and torch.equal(a[1], b[1].cpu()) and torch.equal(a[2], b[2].cpu())
    assert torch.equal(a[3]["num_steps"], b[3]["num_steps"].cpu())
    with pytest.raises(RuntimeError, match="max_steps"):
        gd._integrate(pot, qp, pp, 0.0, 20.0, ts, solver=gd.SemiImplicitEuler(), controller=gd.ConstantStepSize(),
                      dt0=0.1, max_steps=5)


def test_hamiltonian_field_call_forms_reference_doctest():
    """HamiltonianField.__call__ (a-10): the reference's doctest values (orbit/field_hamiltonian.py:108-133) through its
    three array-level call forms."""
    field = gd.HamiltonianField(gp.KeplerPotential(m_tot=1e11))
    x, v = np.array([8.0, 0, 0]), np.array([0, 0.22499668, 0])
    forms = [field(0, x, v, None), field(0, x, v), field(0, (x, v)), field(0, (x, v), None), field(0, np.concatenate([x, v]))]
    for dq, dp in forms:
        assert np.array_equal(dq, v) and np.allclose(dp, [-0.00702891, 0.0, 0.0], rtol=0, atol=5e-9)
    with pytest.raises(NotImplementedError):
        field(0, x, v, {"extra": 1})


def test_time_dependent_potentials_in_stream_release_and_energies():
    """LinearParameter composites beyond the integrators: every stripping time is released in the
    potential of ITS OWN time (df/fardal15.py:49-94 passes t to tidal_radius), and Orbit.total_energy evaluates Phi(q, t)
    at each state's time -- against the oracle frozen at each time."""
    mix = gp.CompositePotential(
        disk=gp.MiyamotoNagaiPotential(gp.LinearParameter(1e7, 0.0, 6.8e10), gp.LinearParameter(1e-4, 100.0, 3.01), 0.28),
        halo=gp.NFWPotential(gp.LinearParameter(2e8, 0.0, 5.4e11), gp.LinearParameter(1e-3, 0.0, 15.62)),
        bulge=gp.HernquistPotential(5e9, 1.0), nuc=gp.JaffePotential(gp.LinearParameter(-1e5, 0.0, 1e9), 0.3),
        iso=gp.IsochronePotential(3e9, gp.LinearParameter(1e-4, 0.0, 2.0)), sat=gp.SatohPotential(2e9, 3.0, gp.LinearParameter(1e-5, 0.0, 0.4)),
        tri=gp.TriaxialHernquistPotential(gp.LinearParameter(1e6, 0.0, 4e9), 0.8, 0.9, gp.LinearParameter(1e-5, 0.0, 0.7)))
    omix = op.Potential((
        op.Component(op.KIND_MN, (6.8e10, 3.0, 0.28), rates=(1e7, 1e-4, 0.0)), op.Component(op.KIND_NFW, (5.4e11, 15.62), rates=(2e8, 1e-3)),
        op.Component(op.KIND_HERNQUIST, (5e9, 1.0)), op.Component(op.KIND_JAFFE, (1e9, 0.3), rates=(-1e5, 0.0)),
        op.Component(op.KIND_ISOCHRONE, (3e9, 2.0), rates=(0.0, 1e-4)), op.Component(op.KIND_SATOH, (2e9, 3.0, 0.4), rates=(0.0, 0.0, 1e-5)),
        op.Component(op.KIND_TRIAXIAL_HERNQUIST, (4e9, 0.8, 0.9, 0.7), rates=(1e6, 0.0, 0.0, 1e-5))))

    def frozen(t):
        return op.Potential(tuple(op.Component(c.kind, c.params_at(t), c.name) for c in omix.components), omix.G)

    rng = np.random.default_rng(51)
    M = 48
    ts = np.linspace(0.0, 2000.0, M)
    xq, xp = synthetic_ics(op.milky_way_potential(), M, seed=52)
    normals = rng.standard_normal((4, M))
    arms = gd.FardalStreamDF().sample(normals, mix, gd.Orbit(q=xq, p=xp, t=ts), 1e4)
    at0 = gd.FardalStreamDF().sample(normals, mix, gd.Orbit(q=xq, p=xp, t=np.zeros(M)), 1e4)
    for i in range(M):
        ql, pl, qt, pt = cref.release_fardal(frozen(ts[i]), xq[i:i + 1], xp[i:i + 1], np.array([1e4]), normals[:, i:i + 1])
        for got, ref in ((arms["lead"].q[i], ql[0]), (arms["lead"].p[i], pl[0]), (arms["trail"].q[i], qt[0]), (arms["trail"].p[i], pt[0])):
            assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max(), i
    assert np.array_equal(at0["lead"].q[0], arms["lead"].q[0])
    assert np.abs(at0["lead"].q[-1] - arms["lead"].q[-1]).max() > 1e-6  # the halo has grown by 74 %: the tidal radius moved
    chen = gd.ChenStreamDF().sample(rng.standard_normal((M, 6)) * 0.1 + np.array([1.6, -30, 0, 1, 20, 0]), mix, gd.Orbit(q=xq, p=xp, t=ts), 1e4)
    assert np.isfinite(chen["lead"].q).all() and np.isfinite(chen["trail"].p).all()

    # energies along an orbit batch: Phi at each save's own time
    q0, p0 = synthetic_ics(op.milky_way_potential(), 32, seed=53)
    tsv = np.linspace(0.0, 800.0, 9)
    orbit = gd.evaluate_orbit(mix, (q0, p0), tsv)
    E = orbit.total_energy()
    ref = np.stack([0.5 * (orbit.p[:, k] ** 2).sum(-1) + op.potential(omix, orbit.q[:, k], tsv[k]) for k in range(9)], axis=1)
    assert E.shape == (32, 9) and np.abs(E / ref - 1).max() < 5e-14
    assert np.abs(gd._energy(mix, orbit.q[:, 3], orbit.p[:, 3], t=tsv[3]) / ref[:, 3] - 1).max() < 5e-14
    assert np.abs(E[:, -1] / E[:, 0] - 1).max() > 1e-3  # not conserved: the potential deepens
    with pytest.raises(NotImplementedError):  # without a time the entry refuses a time-dependent potential
        gd._energy(mix, orbit.q, orbit.p)
    # the whole generator runs in a time-dependent host
    gen = gd.MockStreamGenerator(gd.FardalStreamDF(), mix)
    w0 = gd.PhaseSpaceCoordinate(np.array([30.0, 10, 20]), np.array([10.0, -150, -20]) * KMS, 0.0)
    stream, prog = gen.run(7, np.linspace(0.0, 1500.0, 64), w0, 1e4)
    assert np.isfinite(stream.q).all() and stream.q.shape == (128, 3)


@pytest.mark.parametrize("name", list(PAIRS))
def test_a_particle_does_not_depend_on_the_batch_it_travels_in(name):
    """One arithmetic for every batch size (narrow CTAs for small batches, the same instructions): a particle integrated
    alone, in a batch of 10^3 or inside 10^5 others gives the same bits -- what makes a sharded multi-GPU run equal to
    the unsharded one.  Fixed step (the run-length and the step-by-step kernel, MilkyWayPotential's
    alternating table / closed-form steps included) and Dopri8."""
    cls, ofun = PAIRS[name]
    pot, opot = cls(), ofun()
    q0, p0 = synthetic_ics(opot, 100_000, seed=61)
    ts = np.linspace(0.0, 60.0, 4)
    big = SIE.solve(pot, (q0, p0), 0.0, 60.0, dt0=0.1, saveat=ts)
    for n in (1, 1000):
        small = SIE.solve(pot, (q0[:n], p0[:n]), 0.0, 60.0, dt0=0.1, saveat=ts)
        assert np.array_equal(small.ys[0], big.ys[0][:n]) and np.array_equal(small.ys[1], big.ys[1][:n]), n
    gen = gd._integrate(pot, q0[:1000], p0[:1000], 0.0, 60.0, ts, solver=gd.SemiImplicitEuler(), controller=gd.ConstantStepSize(),
                        dt0=0.1, max_steps=None, general_kernel=True)
    assert np.array_equal(gen[0], big.ys[0][:1000])
    ad = gd.OrbitSolver(stepsize_controller=gd.PIDController(rtol=1e-9, atol=1e-9))
    bigd = ad.solve(pot, (q0[:30_000], p0[:30_000]), 0.0, 200.0, saveat=np.linspace(0, 200.0, 5))
    one = ad.solve(pot, (q0[:257], p0[:257]), 0.0, 200.0, saveat=np.linspace(0, 200.0, 5))
    assert np.array_equal(one.ys[0], bigd.ys[0][:257]) and np.array_equal(one.stats["num_steps"], bigd.stats["num_steps"][:257])


def test_runtime_composites_on_the_combined_table():
    """A composite of the four basic kinds that is none of the three named models (CountsBasicTab): spherical terms in the
    potential's own force table, disks looped over at run time; with at most one disk the steps alternate between table and
    closed forms as MilkyWayPotential's do.  Against the oracle, and run-length == step-by-step kernel bit for bit."""
    one_disk = gp.CompositePotential(disk=gp.MiyamotoNagaiPotential(5e10, 3.0, 0.3), bulge=gp.HernquistPotential(4e9, 0.5),
                                     halo=gp.NFWPotential(8e11, 20.0))
    o_one = op.Potential((op.Component(op.KIND_MN, (5e10, 3.0, 0.3)), op.Component(op.KIND_HERNQUIST, (4e9, 0.5)),
                          op.Component(op.KIND_NFW, (8e11, 20.0))))
    two_disks = gp.CompositePotential(thin=gp.MiyamotoNagaiPotential(4e10, 3.0, 0.25), thick=gp.MiyamotoNagaiPotential(1e10, 2.5, 0.9),
                                      bulge=gp.PowerLawCutoffPotential(5e9, 1.8, 1.9), halo=gp.NFWPotential(6e11, 18.0),
                                      bh=gp.KeplerPotential(4e6))
    o_two = op.Potential((op.Component(op.KIND_MN, (4e10, 3.0, 0.25)), op.Component(op.KIND_MN, (1e10, 2.5, 0.9)),
                          op.Component(op.KIND_PLC, (5e9, 1.8, 1.9)), op.Component(op.KIND_NFW, (6e11, 18.0)),
                          op.Component(op.KIND_HERNQUIST, (4e6, 0.0))))
    for pot, opot in ((one_disk, o_one), (two_disks, o_two)):
        q0, p0 = synthetic_ics(opot, 300, seed=71)
        ts = np.linspace(0.0, 300.0, 4)
        sol = SIE.solve(pot, (q0, p0), 0.0, 300.0, dt0=0.1, saveat=ts)
        gen = gd._integrate(pot, q0, p0, 0.0, 300.0, ts, solver=gd.SemiImplicitEuler(), controller=gd.ConstantStepSize(),
                            dt0=0.1, max_steps=None, general_kernel=True)
        assert np.array_equal(gen[0], sol.ys[0]) and np.array_equal(gen[1], sol.ys[1])
        qr, pr, st, n = cref.integrate_fixed(opot, q0, p0, 0.0, 300.0, 0.1, ts)
        e = rel_dev(sol.ys, (qr, pr))
        assert np.median(e) < 1e-13 and np.mean(e <= 1e-12) > 0.97, (np.median(e), np.mean(e <= 1e-12))
        ad = gd.OrbitSolver(stepsize_controller=gd.PIDController(rtol=1e-10, atol=1e-10)).solve(pot, (q0, p0), 0.0, 300.0, saveat=ts)
        qd, pd, *_ = cref.integrate_dopri8(opot, q0, p0, 0.0, 300.0, ts, rtol=1e-10, atol=1e-10)
        assert np.median(np.abs(ad.ys[0] - qd).max(axis=(1, 2))) < 1e-9


@pytest.mark.parametrize("name", list(PAIRS))
def test_orbits_that_leave_the_force_table(name):
    """The combined spherical table covers r^2 in [2^-8, 2^14) kpc^2 (62 pc .. 128 kpc) for these models; outside it the
    kernels take the closed forms (spherical_fallback).  Orbits that start inside 62 pc, beyond 128 kpc, and that CROSS
    either edge, through every instantiation that looks the table up: the small-launch fixed-step kernel (row fetched
    early with a clamped index), the full-machine one (late loads), the step-by-step kernel and the Dopri8 right-hand
    side -- against the oracle, and bit for bit among themselves."""
    cls, ofun = PAIRS[name]
    pot, opot = cls(), ofun()
    rng = np.random.default_rng(91)
    radii = np.array([0.012, 0.03, 0.055, 0.061, 0.064, 0.08, 0.2, 90.0, 120.0, 127.5, 128.5, 140.0, 250.0, 600.0])
    r = np.repeat(radii, 8)
    u = rng.standard_normal((r.size, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    q0 = r[:, None] * u
    vc = np.sqrt(np.abs(np.sum(q0 * op.gradient(opot, q0), axis=1)))  # circular speed from the oracle's own gradient
    w = rng.standard_normal((r.size, 3)); w /= np.linalg.norm(w, axis=1, keepdims=True)
    p0 = (vc * rng.uniform(0.3, 1.2, r.size))[:, None] * w  # eccentric: most of them cross an edge of the table
    dt, t1 = 0.002, 4.0
    ts = np.array([1.0, 2.5, t1])
    sol = SIE.solve(pot, (q0, p0), 0.0, t1, dt0=dt, saveat=ts)
    qr, pr, st, n = cref.integrate_fixed(opot, q0, p0, 0.0, t1, dt, ts)
    rr = np.linalg.norm(sol.ys[0], axis=-1)
    assert (rr.min() < 0.0625 < rr.max()) and (rr[r > 100].min() < 128.0 < rr.max())  # both edges really are crossed
    e = rel_dev(sol.ys, (qr, pr))
    assert e.max() < 1e-9 and np.median(e) < 1e-12, (e.max(), np.median(e))
    gen = gd._integrate(pot, q0, p0, 0.0, t1, ts, solver=gd.SemiImplicitEuler(), controller=gd.ConstantStepSize(),
                        dt0=dt, max_steps=None, general_kernel=True)
    assert np.array_equal(gen[0], sol.ys[0]) and np.array_equal(gen[1], sol.ys[1])
    # the same particles inside a launch wide enough for the late-load instantiation (more than 512 threads per CTA)
    reps = 90_000 // r.size + 1
    big = SIE.solve(pot, (np.tile(q0, (reps, 1)), np.tile(p0, (reps, 1))), 0.0, t1, dt0=dt, saveat=ts)
    assert np.array_equal(big.ys[0][: r.size], sol.ys[0]) and np.array_equal(big.ys[1][-r.size:], sol.ys[1])
    tol = 1e-10
    ad = gd.OrbitSolver(stepsize_controller=gd.PIDController(rtol=tol, atol=tol))
    a = ad.solve(pot, (q0, p0), 0.0, t1, saveat=ts)
    qd, pd, *_ = cref.integrate_dopri8(opot, q0, p0, 0.0, t1, ts, rtol=tol, atol=tol)
    dq = np.abs(a.ys[0] - qd) / (tol + tol * np.abs(qd))
    dp = np.abs(a.ys[1] - pd) / (tol + tol * np.abs(pd))
    assert dq.max() <= 10.0 and dp.max() <= 10.0, (dq.max(), dp.max())
