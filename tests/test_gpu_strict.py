"""GPU: the reference-order ("strict") fixed-step kernel is the C oracle BIT FOR BIT on all of config C1, and the fast
kernels are held against it -- per particle, next to that particle's own sensitivity to a 1-ulp change of its input.

Why the test has this shape ("prove the fixed-step tail is rounding, not a bug"): the fast kernels differ
from the oracle by <= 1e-12 for 99.4 % of C1's particles and by up to 1e-5 for a handful.  `GX_SCHEME_STRICT` removes
every source of rounding difference (operation order, FMA contraction, MUFU-seeded reciprocals, force tables, libm), so
  (1) strict == oracle exactly: the kernel infrastructure around the arithmetic (time grid, save logic, parameter
      marshalling, summation order) is proven, particle by particle, in all three named models;
  (2) the oracle's own answer moves by `sens_i` when a particle's initial condition is changed by ONE ulp -- measured
      with the strict kernel: median 9e-14, 1 % of C1's orbits above 1.2e-12, the nucleus-scattered ones up to 5e-5.
      Any two correct implementations (XLA:CPU with and without FMA contraction included) differ by rounding errors
      every step, i.e. by at least that;
  (3) the fast kernel's deviation `e_i` is, particle by particle, bounded by that sensitivity: it injects one
      rounding-level perturbation per step, an uncorrelated sequence, so e_i <~ sqrt(n_steps) * sens_i = 100 sens_i
      (measured: median ratio 0.5, 99.9 % below 8.5, max 61), and the north_star bar of 1e-12 holds for every orbit
      on which it can hold for any implementation (100 sens_i <= 1e-12).
There is no hand-picked pericentre mask.
"""
import numpy as np
import pytest

import galax_b200.dynamics as gd
import galax_b200.potential as gp
from oracle import cref
from oracle import potentials as op

from conftest import synthetic_ics

pytestmark = pytest.mark.gpu

PAIRS = {
    "MilkyWayPotential": (gp.MilkyWayPotential, op.milky_way_potential),
    "MilkyWayPotential2022": (gp.MilkyWayPotential2022, op.milky_way_potential_2022),
    "BovyMWPotential2014": (gp.BovyMWPotential2014, op.bovy_mw_potential_2014),
}
FAST = gd.OrbitSolver(solver=gd.SemiImplicitEuler(), stepsize_controller=gd.ConstantStepSize(), max_steps=None)
STRICT = gd.OrbitSolver(solver=gd.SemiImplicitEuler(strict=True), stepsize_controller=gd.ConstantStepSize(),
                        max_steps=None)  # fmt: skip
N_C1 = 10_000


def _dev(a, b):
    (qa, pa), (qb, pb) = a, b
    eq = np.linalg.norm(qa - qb, axis=-1) / np.linalg.norm(qb, axis=-1)
    ep = np.linalg.norm(pa - pb, axis=-1) / np.linalg.norm(pb, axis=-1)
    return np.maximum(eq, ep).max(axis=-1)


def _one_ulp(x, rng):
    up = rng.integers(0, 2, size=x.shape).astype(bool)
    return np.where(up, np.nextafter(x, np.inf), np.nextafter(x, -np.inf))


@pytest.mark.parametrize("name", list(PAIRS))
def test_c1_strict_is_the_oracle_bit_for_bit_and_bounds_the_fast_kernel(name):
    """Config C1 exactly: 10^4 particles, dt = 0.1 Myr over 1 Gyr = 10^4 steps."""
    cls, ofun = PAIRS[name]
    pot, opot = cls(), ofun()
    q0, p0 = synthetic_ics(opot, N_C1, seed=1)
    s = STRICT.solve(pot, (q0, p0), 0.0, 1000.0, dt0=0.1)
    qr, pr, st, n = cref.integrate_fixed(opot, q0, p0, 0.0, 1000.0, 0.1, [1000.0])
    assert (n == 10000).all() and (st == 0).all()
    # (1) bit for bit, every particle
    assert np.array_equal(s.ys[0], qr) and np.array_equal(s.ys[1], pr)

    # (2) each orbit's own sensitivity to one ulp of its initial condition (max over 4 sign patterns)
    rng = np.random.default_rng(7)
    sens = np.zeros(N_C1)
    for _ in range(4):
        sj = STRICT.solve(pot, (_one_ulp(q0, rng), _one_ulp(p0, rng)), 0.0, 1000.0, dt0=0.1)
        sens = np.maximum(sens, _dev(sj.ys, s.ys))

    # (3) the fast kernel against the reference-order result
    f = FAST.solve(pot, (q0, p0), 0.0, 1000.0, dt0=0.1)
    e = _dev(f.ys, (qr, pr))
    assert (e <= np.maximum(1e-12, 100.0 * sens)).all(), (e / np.maximum(sens, 1e-17)).max()
    # ... and as distributions: closer to the oracle than the oracle is to itself one ulp away
    assert np.median(e) <= np.median(sens) and np.quantile(e, 0.99) <= np.quantile(sens, 0.99)
    assert np.median(e) <= 1e-13 and np.quantile(e, 0.99) <= 1.5e-12
    assert np.mean(e <= 1e-12) >= np.mean(sens <= 1e-12) - 0.005


def test_strict_saves_interpolation_backward_and_leapfrog_midpoint_bitwise():
    """The rest of the strict kernel's surface against the oracle, still bit for bit: saves on and between step
    boundaries, several saves per step, a clipped last step, backward integration, LeapfrogMidpoint, max_steps."""
    pot, opot = gp.MilkyWayPotential2022(), op.milky_way_potential_2022()
    q0, p0 = synthetic_ics(opot, 333, seed=5)
    ts = np.concatenate([[0.0], np.sort(np.random.default_rng(0).uniform(0, 77.77, 41)), [77.77]])
    s = STRICT.solve(pot, (q0, p0), 0.0, 77.77, saveat=ts, dt0=0.3)
    qr, pr, st, n = cref.integrate_fixed(opot, q0, p0, 0.0, 77.77, 0.3, ts)
    assert np.array_equal(s.ys[0], qr) and np.array_equal(s.ys[1], pr)
    s = STRICT.solve(pot, (q0, p0), 0.0, -50.0, dt0=-0.1)
    qr, pr, st, n = cref.integrate_fixed(opot, q0, p0, 0.0, -50.0, -0.1, [-50.0])
    assert np.array_equal(s.ys[0], qr) and np.array_equal(s.ys[1], pr)
    lfm = gd.OrbitSolver(solver=gd.LeapfrogMidpoint(strict=True), stepsize_controller=gd.ConstantStepSize(), max_steps=None)
    s = lfm.solve(pot, (q0, p0), 0.0, 100.0, saveat=np.linspace(0, 100, 7), dt0=0.05)
    qr, pr, st, n = cref.integrate_fixed(opot, q0, p0, 0.0, 100.0, 0.05, np.linspace(0, 100, 7), scheme=1)
    assert np.array_equal(s.ys[0], qr) and np.array_equal(s.ys[1], pr)
    s = STRICT.solve(pot, (q0, p0), 0.0, 100.0, dt0=0.05, max_steps=10, throw=False)
    assert (np.asarray(s.result) == 1).all() and np.isnan(s.ys[0]).all()
    # the T3N layout is the same numbers
    q2, p2, _, _ = gd._integrate(pot, q0, p0, 0.0, 77.77, ts, solver=gd.SemiImplicitEuler(strict=True),
                                 controller=gd.ConstantStepSize(), dt0=0.3, max_steps=None, layout="T3N")
    s = STRICT.solve(pot, (q0, p0), 0.0, 77.77, saveat=ts, dt0=0.3)
    assert np.array_equal(np.transpose(q2, (2, 0, 1)), s.ys[0]) and np.array_equal(np.transpose(p2, (2, 0, 1)), s.ys[1])


def _tolunits_t(a, ref, tol):
    """|a - ref| / (atol + rtol |ref|), max over the six components: [N, T]."""
    d = 0.0
    for x, r in zip(a, ref):
        d = np.maximum(d, (np.abs(x - r) / (tol + tol * np.abs(r))).max(axis=2))
    return d


def test_c2_dopri8_strict_is_the_oracle_bit_for_bit_and_bounds_the_fast_kernel():
    """Config C2's shape on its 4096-particle parity subset: MilkyWayPotential2022, Dopri8 + PID with
    rtol = atol = 1e-10, dt0 = None (initial-step heuristic), 1000 saves over 5 Gyr.

    (1) The reference-order kernel (`GX_SOLVER_STRICT`) equals oracle/galax_oracle.c BIT FOR BIT at all 1000 saves of
        every particle, with the same accepted / attempted step counts: the same step sequence for 100 % of the
        particles, the north_star bar (10 x tol) met with zero deviation.
    (2) What that bar can mean for an implementation that rounds differently: the strict kernel itself, started one
        ulp away, leaves the bar for 77 % of the particles within 5 Gyr (median 43 tolerance units, measured) -- the
        controller's accept / reject decisions sit on the last bits of the error estimate, and once two runs take
        different steps they differ by the method's own global / dense-output error.
    (3) The fast kernel against the oracle, save time by save time, is held to that twin: at most 4 x its deviation in
        the median and in the 99th percentile; the median particle is inside 10 x tol at EVERY save time; the step
        counts agree in total to 0.1 %; energy is conserved to 1e3 x tol."""
    tol = 1e-10
    pot, opot = gp.MilkyWayPotential2022(), op.milky_way_potential_2022()
    q0, p0 = synthetic_ics(opot, 4096, seed=2)
    ts = np.linspace(0.0, 5000.0, 1000)
    ctl = gd.PIDController(rtol=tol, atol=tol)
    strict = gd.OrbitSolver(solver=gd.Dopri8(strict=True), stepsize_controller=ctl, max_steps=2**16)
    fast = gd.OrbitSolver(solver=gd.Dopri8(), stepsize_controller=ctl, max_steps=2**16)
    qr, pr, st, na, nt = cref.integrate_dopri8(opot, q0, p0, 0.0, 5000.0, ts, rtol=tol, atol=tol, max_steps=2**16)
    assert (st == 0).all()
    s = strict.solve(pot, (q0, p0), 0.0, 5000.0, saveat=ts)
    assert np.array_equal(s.ys[0], qr) and np.array_equal(s.ys[1], pr)  # (1)
    assert np.array_equal(np.asarray(s.stats["num_steps"]), nt) and np.array_equal(np.asarray(s.stats["num_accepted_steps"]), na)

    rng = np.random.default_rng(7)
    twin = 0.0
    for _ in range(2):  # (2)
        sj = strict.solve(pot, (_one_ulp(q0, rng), _one_ulp(p0, rng)), 0.0, 5000.0, saveat=ts)
        twin = np.maximum(twin, _tolunits_t(sj.ys, (qr, pr), tol))
    assert np.mean(twin.max(axis=1) <= 10.0) < 0.6  # the bar is not reachable by rounding differently: documented fact

    f = fast.solve(pot, (q0, p0), 0.0, 5000.0, saveat=ts)  # (3)
    e = _tolunits_t(f.ys, (qr, pr), tol)
    assert np.median(e, axis=0).max() <= 10.0
    for k in (1, 10, 100, 500, 999):
        assert np.median(e[:, k]) <= 4.0 * np.median(twin[:, k]) + 0.2, k
        assert np.quantile(e[:, k], 0.99) <= 4.0 * np.quantile(twin[:, k], 0.99) + 2.0, k
    assert np.mean(e.max(axis=1) <= 10.0) >= np.mean(twin.max(axis=1) <= 10.0) - 0.25
    assert abs(int(np.asarray(f.stats["num_steps"]).sum()) / int(nt.sum()) - 1) < 1e-3
    E0, E1 = gd._energy(pot, q0, p0), gd._energy(pot, f.ys[0][:, -1], f.ys[1][:, -1])
    assert np.quantile(np.abs(E1 / E0 - 1), 0.99) < 1e3 * tol


@pytest.mark.parametrize("name", list(PAIRS))
def test_strict_dopri_all_models_dopri5_backward_and_per_particle_t0(name):
    """The rest of the strict adaptive kernel's surface, bit for bit against the oracle: the three named models at two
    tolerances with a given first step, Dopri5, backward integration, per-particle start times, forced dtmin."""
    cls, ofun = PAIRS[name]
    pot, opot = cls(), ofun()
    q0, p0 = synthetic_ics(opot, 200, seed=21)
    ts = np.linspace(0.0, 400.0, 9)
    for tol, dt0 in ((1e-7, None), (1e-11, 0.5)):
        sol = gd.OrbitSolver(solver=gd.Dopri8(strict=True), stepsize_controller=gd.PIDController(tol, tol)).solve(
            pot, (q0, p0), 0.0, 400.0, saveat=ts, dt0=dt0)
        qr, pr, st, na, nt = cref.integrate_dopri8(opot, q0, p0, 0.0, 400.0, ts, rtol=tol, atol=tol, max_steps=2**16,
                                                   **({} if dt0 is None else {"dt0": dt0}))
        assert np.array_equal(sol.ys[0], qr) and np.array_equal(sol.ys[1], pr)
        assert np.array_equal(np.asarray(sol.stats["num_steps"]), nt)
    sol = gd.OrbitSolver(solver=gd.Dopri5(strict=True), stepsize_controller=gd.PIDController(1e-7, 1e-7, dtmin=0.3)).solve(
        pot, (q0, p0), 0.0, -300.0, saveat=np.linspace(0.0, -300.0, 5))
    qr, pr, st, na, nt = cref.integrate_dopri8(opot, q0, p0, 0.0, -300.0, np.linspace(0.0, -300.0, 5), rtol=1e-7, atol=1e-7,
                                               solver="dopri5", dtmin=0.3, max_steps=2**16)
    assert np.array_equal(sol.ys[0], qr) and np.array_equal(sol.ys[1], pr)
    t0 = np.random.default_rng(3).uniform(0.0, 350.0, 200)
    q, p, status, stats = gd._integrate(pot, q0, p0, t0, 400.0, np.array([400.0]), solver=gd.Dopri8(strict=True),
                                        controller=gd.PIDController(1e-8, 1e-8), dt0=None, max_steps=None)
    qr, pr, st, na, nt = cref.integrate_dopri8(opot, q0, p0, t0, 400.0, [400.0], rtol=1e-8, atol=1e-8)
    assert np.array_equal(q, qr) and np.array_equal(p, pr)


def test_strict_composites_and_refusals():
    """Composites built from the four basic kinds run (a lone MN3 disk, a user composite); anything else is refused."""
    mn3, omn3 = gp.MN3Sech2Potential(m_tot=4.7717e10, h_R=2.6, h_z=0.3, positive_density=True), op.mn3_potential(
        4.7717e10, 2.6, 0.3, sech2=True, positive_density=True)
    q0, p0 = synthetic_ics(op.milky_way_potential(), 64, seed=9)
    s = STRICT.solve(mn3, (q0, p0), 0.0, 30.0, dt0=0.1)
    qr, pr, _, _ = cref.integrate_fixed(omn3, q0, p0, 0.0, 30.0, 0.1, [30.0])
    assert np.array_equal(s.ys[0], qr) and np.array_equal(s.ys[1], pr)
    with pytest.raises(Exception, match="unsupported"):
        STRICT.solve(gp.LM10Potential(), (q0, p0), 0.0, 1.0, dt0=0.1)
