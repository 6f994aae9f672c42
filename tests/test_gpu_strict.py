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
