"""GPU: Orbit post-processing fused into the integrator kernels (gx_integrate_*_epilogue): total energy,
angular momentum and tidal tensor at every saved state, against the stand-alone device pass, the numpy oracle
(coordinates/_src/pscs/base.py:182-330, potential/_src/register_funcs.py:347-377) and the plain integrators."""
import numpy as np
import pytest

import galax_b200.dynamics as gd
import galax_b200.potential as gp
from oracle import potentials as op

from conftest import synthetic_ics

pytestmark = pytest.mark.gpu

ALL = ("energy", "angular_momentum", "tidal_tensor")
PAIRS = {
    "MilkyWayPotential": (gp.MilkyWayPotential, op.milky_way_potential),
    "MilkyWayPotential2022": (gp.MilkyWayPotential2022, op.milky_way_potential_2022),
    "BovyMWPotential2014": (gp.BovyMWPotential2014, op.bovy_mw_potential_2014),
}
SIE = dict(solver=gd.SemiImplicitEuler(), controller=gd.ConstantStepSize(), dt0=0.5, max_steps=None)
DP8 = dict(solver=gd.Dopri8(), controller=gd.PIDController(1e-9, 1e-9), dt0=None, max_steps=2**16)


def oracle_diag(opot, q, p):
    qf, pf = q.reshape(-1, 3), p.reshape(-1, 3)
    E = 0.5 * (pf**2).sum(-1) + op.potential(opot, qf)
    L = np.cross(qf, pf)
    H = op.hessian(opot, qf)
    TT = H - np.trace(H, axis1=-2, axis2=-1)[:, None, None] * np.eye(3) / 3.0
    return E.reshape(q.shape[:-1]), L.reshape(q.shape), TT.reshape(q.shape[:-1] + (3, 3))


def check(pot, opot, q, p, st, plain=None):
    if plain is not None:  # the epilogue kernels integrate exactly like the plain ones
        assert np.array_equal(q, plain[0]) and np.array_equal(p, plain[1])
    E, L, TT = oracle_diag(opot, q, p)
    assert np.abs(st["energy"] / E - 1).max() < 2e-14
    assert np.abs(st["angular_momentum"] - L).max() <= 4e-16 * np.abs(L).max()
    assert np.abs(st["tidal_tensor"] - TT).max() <= 2e-13 * np.abs(TT).max()
    # ... and the stand-alone passes over the saved states give the same bits
    assert np.array_equal(st["energy"], gd._energy(pot, q, p))
    assert np.array_equal(st["angular_momentum"], gd._energy(None, q, p, want="L"))
    assert np.abs(st["tidal_tensor"] - pot.tidal_tensor(q.reshape(-1, 3)).reshape(TT.shape)).max() <= 1e-15 * np.abs(TT).max()


@pytest.mark.parametrize("name", list(PAIRS))
@pytest.mark.parametrize("T", [1, 3, 4, 37])
def test_fixed_step_epilogue(name, T):
    """rows shorter than 4 saves are stored directly, longer ones staged per lane and flushed in whole sectors; the
    particle count is not a multiple of 4 either, so the rows start at every sector phase"""
    cls, ofun = PAIRS[name]
    pot, opot = cls(), ofun()
    q0, p0 = synthetic_ics(opot, 333, seed=11)
    ts = np.linspace(0.0, 200.0, T) if T > 1 else np.array([200.0])
    q, p, status, st = gd._integrate(pot, q0, p0, 0.0, 200.0, ts, diagnostics=ALL, **SIE)
    plain = gd._integrate(pot, q0, p0, 0.0, 200.0, ts, **SIE)
    assert st["energy"].shape == (333, T) and st["angular_momentum"].shape == (333, T, 3)
    assert st["tidal_tensor"].shape == (333, T, 3, 3)
    check(pot, opot, q, p, st, plain)


def test_fixed_step_epilogue_other_kernels_and_layouts():
    pot, opot = gp.MilkyWayPotential(), op.milky_way_potential()
    q0, p0 = synthetic_ics(opot, 200, seed=12)
    ts = np.linspace(0.0, 100.0, 21)
    ref = gd._integrate(pot, q0, p0, 0.0, 100.0, ts, diagnostics=ALL, **SIE)
    # the general (step-by-step) kernel and the structure-of-arrays layout
    g = gd._integrate(pot, q0, p0, 0.0, 100.0, ts, diagnostics=ALL, general_kernel=True, **SIE)
    for d in ALL:
        assert np.array_equal(g[3][d], ref[3][d])
    t3n = gd._integrate(pot, q0, p0, 0.0, 100.0, ts, diagnostics=ALL, layout="T3N", **SIE)
    assert np.array_equal(np.asarray(t3n[3]["energy"]).T, ref[3]["energy"])
    assert np.array_equal(np.asarray(t3n[3]["angular_momentum"]).transpose(2, 0, 1), ref[3]["angular_momentum"])
    assert np.array_equal(np.asarray(t3n[3]["tidal_tensor"]).transpose(2, 0, 1).reshape(200, 21, 3, 3), ref[3]["tidal_tensor"])
    # only some of the outputs; LeapfrogMidpoint; a large batch (wide CTAs, no table variant)
    e = gd._integrate(pot, q0, p0, 0.0, 100.0, ts, diagnostics=("energy",), **SIE)
    assert set(e[3]) == {"energy"} and np.array_equal(e[3]["energy"], ref[3]["energy"])
    l = gd._integrate(pot, q0, p0, 0.0, 100.0, ts, diagnostics=("angular_momentum",), **SIE)
    assert np.array_equal(l[3]["angular_momentum"], ref[3]["angular_momentum"])
    lf = dict(SIE, solver=gd.LeapfrogMidpoint())
    q, p, _, st = gd._integrate(pot, q0, p0, 0.0, 100.0, ts, diagnostics=ALL, **lf)
    check(pot, opot, q, p, st, gd._integrate(pot, q0, p0, 0.0, 100.0, ts, **lf))
    qb, pb = synthetic_ics(opot, 148 * 128 * 4 + 5, seed=13)
    q, p, _, st = gd._integrate(pot, qb, pb, 0.0, 20.0, np.linspace(0, 20.0, 6), diagnostics=("energy", "angular_momentum"), **SIE)
    E, L, _ = oracle_diag(opot, q[:4096], p[:4096])
    assert np.abs(st["energy"][:4096] / E - 1).max() < 2e-14 and np.abs(st["angular_momentum"][:4096] - L).max() <= 4e-16 * np.abs(L).max()


@pytest.mark.parametrize("name", list(PAIRS))
@pytest.mark.parametrize("T", [2, 10, 101])
def test_dopri8_epilogue(name, T):
    cls, ofun = PAIRS[name]
    pot, opot = cls(), ofun()
    q0, p0 = synthetic_ics(opot, 301, seed=14)
    ts = np.linspace(0.0, 500.0, T)
    q, p, status, st = gd._integrate(pot, q0, p0, 0.0, 500.0, ts, diagnostics=ALL, fuse=True, **DP8)
    plain = gd._integrate(pot, q0, p0, 0.0, 500.0, ts, **DP8)
    check(pot, opot, q, p, st, plain)
    # the default policy (fuse="auto": a second pass beyond 16 saves in the adaptive kernel) returns the same arrays
    auto = gd._integrate(pot, q0, p0, 0.0, 500.0, ts, diagnostics=ALL, **DP8)[3]
    assert np.array_equal(auto["energy"], st["energy"]) and np.array_equal(auto["angular_momentum"], st["angular_momentum"])
    assert np.abs(auto["tidal_tensor"] - st["tidal_tensor"]).max() <= 1e-15 * np.abs(st["tidal_tensor"]).max()
    assert np.array_equal(st["num_steps"], plain[3]["num_steps"])
    # energy is conserved to the tolerance along every orbit, the z angular momentum exactly to rounding (axisymmetric)
    E = st["energy"]
    assert np.abs(E / E[:, :1] - 1).max() < 1e-6
    Lz = st["angular_momentum"][..., 2]
    assert np.abs(Lz - Lz[:, :1]).max() <= 1e-7 * np.abs(Lz).max()


def test_dopri8_epilogue_unreached_saves_are_nan_and_runtime_composites():
    pot, opot = gp.MilkyWayPotential(), op.milky_way_potential()
    q0, p0 = synthetic_ics(opot, 64, seed=15)
    ts = np.linspace(0.0, 3000.0, 40)
    kw = dict(DP8, max_steps=60)
    q, p, status, st = gd._integrate(pot, q0, p0, 0.0, 3000.0, ts, diagnostics=ALL, throw=False, fuse=True, **kw)
    assert (np.asarray(status) == 1).any()
    nanq = np.isnan(q[..., 0])
    assert nanq.any() and np.array_equal(np.isnan(st["energy"]), nanq)
    assert np.array_equal(np.isnan(st["angular_momentum"][..., 1]), nanq) and np.array_equal(np.isnan(st["tidal_tensor"][..., 2, 2]), nanq)
    ok = ~nanq
    E, L, TT = oracle_diag(opot, q[ok], p[ok])
    assert np.abs(st["energy"][ok] / E - 1).max() < 2e-14
    # a runtime composite (not one of the three named models) and per-particle start times
    comp = gp.CompositePotential(disk=gp.MiyamotoNagaiPotential(m_tot=5e10, a=3.0, b=0.3), halo=gp.NFWPotential(m=6e11, r_s=18.0),
                                 bulge=gp.HernquistPotential(m_tot=4e9, r_s=0.8), extra=gp.PlummerPotential(m_tot=1e9, r_s=2.0))
    t0 = np.linspace(0.0, 100.0, 64)
    tsb = np.linspace(100.0, 400.0, 9)
    q, p, status, st = gd._integrate(comp, q0, p0, t0, 400.0, tsb, diagnostics=("energy", "angular_momentum"), **DP8)
    # (a runtime composite: the fused epilogue and the stand-alone pass are different instantiations of the evaluator,
    #  so rounding-level, not bit-level, agreement)
    assert np.abs(st["energy"] / gd._energy(comp, q, p) - 1).max() < 1e-15 and np.isfinite(st["energy"]).all()


def test_epilogue_refuses_what_it_cannot_do():
    pot, opot = gp.MilkyWayPotential(), op.milky_way_potential()
    q0, p0 = synthetic_ics(opot, 8, seed=16)
    with pytest.raises(ValueError):
        gd._integrate(pot, q0, p0, 0.0, 10.0, np.array([10.0]), diagnostics=("entropy",), **SIE)
    with pytest.raises(NotImplementedError):  # reference-order kernels: no epilogue
        gd._integrate(pot, q0, p0, 0.0, 10.0, np.array([10.0]), diagnostics=("energy",), **dict(SIE, solver=gd.SemiImplicitEuler(strict=True)))
    td = gp.HernquistPotential(m_tot=gp.LinearParameter(slope=1e6, point_time=0.0, point_value=1e10), r_s=1.0)
    with pytest.raises(NotImplementedError):  # E = |p|^2/2 + Phi(q, t): static potentials only
        gd._integrate(td, q0, p0, 0.0, 10.0, np.array([10.0]), diagnostics=("energy",), **DP8)


def test_orbit_api_returns_the_fused_diagnostics():
    pot, opot = gp.MilkyWayPotential2022(), op.milky_way_potential_2022()
    q0, p0 = synthetic_ics(opot, 50, seed=17)
    t = np.linspace(0.0, 300.0, 13)
    orbit = gd.evaluate_orbit(pot, (q0, p0), t, diagnostics=ALL)
    assert orbit.total_energy() is orbit.diagnostics["energy"] and orbit.diagnostics["energy"].shape == (50, 13)
    assert orbit.angular_momentum() is orbit.diagnostics["angular_momentum"]
    assert orbit.tidal_tensor() is orbit.diagnostics["tidal_tensor"]
    plain = gd.evaluate_orbit(pot, (q0, p0), t)
    assert plain.diagnostics is None and np.array_equal(plain.q, orbit.q)
    assert np.array_equal(plain.total_energy(), orbit.total_energy())  # (the second pass: same bits)
    assert np.abs(plain.tidal_tensor() - orbit.tidal_tensor()).max() <= 1e-15 * np.abs(orbit.tidal_tensor()).max()
    other = gp.MilkyWayPotential()
    assert not np.array_equal(orbit.total_energy(other), orbit.total_energy())  # another potential: recomputed
    co = gd.compute_orbit(pot, (q0, p0), t, diagnostics=("energy",))
    assert co.total_energy() is co.diagnostics["energy"]
