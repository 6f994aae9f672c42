"""The reference-side binding (galax_b200/interop/galax_plugin.py) exercised without galax: duck-typed stand-ins for
galax / unxt / plum / jax / coordinax / diffrax (tests/fake_galax, same public attribute names and call forms as the
reference, cited there) drive ``convert_potential`` and every registered overload.

CPU: the marshalling -- a galax object becomes the same ``gx_potential`` bytes as the native constructor -- and the
registration itself.  GPU: the overloads return the reference's container types with the kernels' numbers."""
import ctypes
import importlib
import sys
from pathlib import Path

import numpy as np
import pytest

FAKE = str(Path(__file__).resolve().parent / "fake_galax")
FAKE_MODULES = ("galax", "galax.potential", "galax.dynamics", "galax.coordinates", "unxt", "plum", "jax", "jax.numpy",
                "coordinax", "coordinax.vecs", "diffrax")  # fmt: skip


@pytest.fixture()
def plugin():
    """Import the plugin with the stand-ins on sys.path; undo everything afterwards."""
    saved = {m: sys.modules.pop(m) for m in list(sys.modules) if m.split(".")[0] in {x.split(".")[0] for x in FAKE_MODULES}}
    sys.path.insert(0, FAKE)
    try:
        import galax_b200.interop.galax_plugin as plug

        plug.__dict__.pop("_IMPORT_ERROR", None)  # (reload keeps the old namespace)
        plug = importlib.reload(plug)
        assert not hasattr(plug, "_IMPORT_ERROR"), getattr(plug, "_IMPORT_ERROR", None)
        yield plug
    finally:
        sys.path.remove(FAKE)
        for m in list(sys.modules):
            if m.split(".")[0] in {x.split(".")[0] for x in FAKE_MODULES}:
                del sys.modules[m]
        sys.modules.update(saved)
        import galax_b200.interop.galax_plugin as plug

        importlib.reload(plug)  # inert again (galax not importable)


def _bytes(pot):
    P = pot.c_struct()
    return ctypes.string_at(ctypes.byref(P), ctypes.sizeof(P))


def test_convert_potential_gives_the_native_bytes(plugin):
    import galax.potential as gp
    import unxt as u

    import galax_b200.potential as bp

    # the three named models (+ LM10), parameters read from the live objects; lengths given in pc arrive in kpc
    for fake, native in ((gp.MilkyWayPotential(), bp.MilkyWayPotential()), (gp.MilkyWayPotential2022(), bp.MilkyWayPotential2022()),
                         (gp.BovyMWPotential2014(), bp.BovyMWPotential2014()), (gp.LM10Potential(), bp.LM10Potential())):  # fmt: skip
        got = plugin.convert_potential(fake)
        assert _bytes(got) == _bytes(native), type(fake).__name__
    # user-overridden parameters and G propagate
    mw = gp.MilkyWayPotential(disk=gp.MiyamotoNagaiPotential(m_tot=u.Q(5e10, "Msun"), a=u.Q(2500.0, "pc"), b=0.3))
    mw.constants["G"].value = 4.3e-12
    try:
        got = plugin.convert_potential(mw)
        assert got.G == 4.3e-12
        assert _bytes(got) == _bytes(bp.MilkyWayPotential(disk=dict(m_tot=5e10, a=2.5, b=0.3), G=4.3e-12))
    finally:
        mw.constants["G"].__class__.value = gp.G_GALACTIC
    # single components, an MN3 disk on its own (host-side fit redone with the same operation order), a user composite
    assert _bytes(plugin.convert_potential(gp.MN3Sech2Potential(m_tot=1e11, h_R=3.0, h_z=0.4))) == _bytes(
        bp.MN3Sech2Potential(1e11, 3.0, 0.4))
    comp = gp.CompositePotential({"a": gp.HernquistPotential(m_tot=1e10, r_s=1.0), "b": gp.KeplerPotential(m_tot=1e9),
                                  "c": gp.PlummerPotential(m_tot=1e8, r_s=0.5)})  # fmt: skip
    nat = bp.CompositePotential({"a": bp.HernquistPotential(1e10, 1.0), "b": bp.KeplerPotential(1e9), "c": bp.PlummerPotential(1e8, 0.5)})
    assert _bytes(plugin.convert_potential(comp)) == _bytes(nat)
    # LinearParameter: the reference's doctest potential (Kepler losing 1e3 Msun / yr): slope in Msun / Myr, offset folded
    lp = gp.LinearParameter(slope=u.Q(-1e3, "Msun / yr"), point_time=u.Q(0.0, "Myr"), point_value=u.Q(1e12, "Msun"))
    got = plugin.convert_potential(gp.KeplerPotential(m_tot=lp))
    want = bp.KeplerPotential(m_tot=bp.LinearParameter(slope=-1e9, point_time=0.0, point_value=1e12))
    assert _bytes(got) == _bytes(want) and got.is_time_dependent
    # anything else is refused, loudly
    with pytest.raises(NotImplementedError, match="ConstantParameter and LinearParameter"):
        plugin.convert_potential(gp.HernquistPotential(m_tot=gp.UserParameter(lambda t: 1e10), r_s=1.0))
    with pytest.raises(NotImplementedError, match="no galax_b200 kernel"):
        plugin.convert_potential(type("BarPotential", (gp.AbstractPotential,), {})())


def test_every_reference_call_form_has_an_overload(plugin):
    import plum

    counts = {k: len(v) for k, v in plum.REGISTRY.items()}
    # register_funcs.py:86-155 (six forms), :276-320 (four), legacy/funcs.py:42-51 + :216-254, orbit/compute.py:28-98
    assert counts == {"gradient": 6, "hessian": 4, "acceleration": 1, "tidal_tensor": 1, "evaluate_orbit": 2,
                      "compute_orbit": 1}  # fmt: skip
    assert plugin.MockStreamGenerator.run.__code__.co_varnames[:5] == ("self", "rng", "ts", "prog_w0", "prog_mass")


def test_solver_records_are_translated(plugin):
    import diffrax as dfx
    import galax.dynamics as gd

    import galax_b200.dynamics as bd

    spec = plugin._integrator_spec(gd.Integrator(
        dynamics_solver=gd.OrbitSolver(solver=dfx.Dopri8(), stepsize_controller=dfx.PIDController(
            rtol=1e-9, atol=1e-11, pcoeff=0.1, icoeff=0.8, dcoeff=0.05, dtmin=0.05, factormax=5.0), max_steps=1234),
        diffeq_kw={"max_steps": None, "dt0": 0.5}))  # fmt: skip
    c = spec.dynamics_solver.stepsize_controller
    assert isinstance(spec.dynamics_solver.solver, bd.Dopri8) and spec.dynamics_solver.max_steps == 1234
    assert (c.rtol, c.atol, c.pcoeff, c.icoeff, c.dcoeff, c.dtmin, c.factormax) == (1e-9, 1e-11, 0.1, 0.8, 0.05, 0.05, 5.0)
    assert spec.diffeq_kw == {"max_steps": None, "dt0": 0.5}
    s = plugin._solver_spec(gd.OrbitSolver(solver=dfx.SemiImplicitEuler(), stepsize_controller=dfx.ConstantStepSize()))
    assert isinstance(s.solver, bd.SemiImplicitEuler) and isinstance(s.stepsize_controller, bd.ConstantStepSize)
    assert isinstance(plugin._solver_spec(gd.OrbitSolver(solver=dfx.Dopri5())).solver, bd.Dopri5)
    assert isinstance(plugin._integrator_spec(None).dynamics_solver.solver, bd.Dopri8)
    with pytest.raises(NotImplementedError):
        plugin._solver_spec(gd.OrbitSolver(solver=dfx.SemiImplicitEuler()))  # SIE needs ConstantStepSize


@pytest.mark.gpu
def test_overloads_return_reference_containers_with_kernel_numbers(plugin):
    import coordinax as cx
    import diffrax as dfx
    import galax.coordinates as gc
    import galax.dynamics as gd
    import galax.potential as gp
    import plum
    import unxt as u

    import galax_b200.dynamics as bd
    import galax_b200.potential as bp

    G = {k: plum._Generic(k) for k in plum.REGISTRY}
    pot, nat = gp.MilkyWayPotential(), bp.MilkyWayPotential()
    x = np.array([[1.0, 2.0, 3.0], [8.0, 0.5, -0.3]])
    g_ref, h_ref = nat.gradient(x), nat.hessian(x)
    # the reference's known-answer point (tests/unit/potential/builtin/test_milkywaypotential.py:41-70)
    assert np.allclose(g_ref[0], [0.00256407, 0.00512815, 0.01115285], atol=1e-8)
    # arrays in -> bare arrays out (positional and t= forms)
    assert np.array_equal(G["gradient"](pot, x, 0.0), g_ref) and np.array_equal(G["gradient"](pot, x, t=0.0), g_ref)
    assert np.array_equal(G["hessian"](pot, x, 0.0), h_ref) and np.array_equal(G["hessian"](pot, x, t=0.0), h_ref)
    assert np.array_equal(G["acceleration"](pot, x, 0.0), -g_ref)
    tt = G["tidal_tensor"](pot, x, 0.0)
    assert np.allclose(tt, h_ref - np.trace(h_ref, axis1=-2, axis2=-1)[:, None, None] * np.eye(3) / 3, rtol=0, atol=1e-18)
    # Quantities in (pc!) -> Quantities out in the potential's units
    gq = G["gradient"](pot, u.Q(x * 1e3, "pc"), u.Q(0.0, "Gyr"))
    assert isinstance(gq, u.Quantity) and gq.unit is pot.units["acceleration"] and np.allclose(gq.value, g_ref, rtol=1e-15)
    gq = G["gradient"](pot, u.Q(x, "kpc"), t=u.Q(0.0, "Myr"))
    assert isinstance(gq, u.Quantity) and np.array_equal(gq.value, g_ref)
    hq = G["hessian"](pot, u.Q(x, "kpc"), u.Q(0.0, "Myr"))
    assert isinstance(hq, u.Quantity) and hq.unit is pot.units["frequency drift"] and np.array_equal(hq.value, h_ref)
    # vectors / phase-space objects in -> CartesianAcc3D out
    gv = G["gradient"](pot, cx.vecs.CartesianPos3D.from_(x, "kpc"), t=u.Q(0.0, "Myr"))
    assert isinstance(gv, cx.vecs.CartesianAcc3D) and np.array_equal(gv.xyz.value, g_ref)
    w = gc.PhaseSpaceCoordinate(q=u.Q(x, "kpc"), p=u.Q(np.zeros_like(x), "km / s"), t=u.Q(0.0, "Gyr"))
    assert np.array_equal(G["gradient"](pot, w).xyz.value, g_ref)

    # evaluate_orbit: phase-space object with units (km/s), Quantity times (Gyr); Orbit back, frame kept
    q0 = np.array([[8.0, 0.0, 0.0], [10.0, 1.0, 2.0]])
    p0_kms = np.array([[0.0, 220.0, 0.0], [30.0, 180.0, 20.0]])
    w0 = gc.PhaseSpaceCoordinate(q=u.Q(q0, "kpc"), p=u.Q(p0_kms, "km / s"), t=u.Q(0.0, "Gyr"))
    ts = u.Q(np.linspace(0.0, 500.0, 6), "Myr")  # (Gyr works too; Myr keeps the save times bit-identical to the native call)
    orb = G["evaluate_orbit"](pot, w0, ts)
    ref = bd.evaluate_orbit(nat, bd.PhaseSpaceCoordinate(q0, p0_kms * u.KMS, 0.0), np.linspace(0.0, 500.0, 6))
    assert isinstance(orb, gd.Orbit) and orb.frame == w0.frame and orb.q.unit is pot.units["length"]
    assert np.array_equal(orb.q.value, ref.q) and np.array_equal(orb.p.value, ref.p) and np.array_equal(orb.t.value, ref.t)
    orb2 = G["evaluate_orbit"](pot, w0, t=ts)
    assert np.array_equal(orb2.q.value, ref.q)
    # ... with the reference's fixed-step "leapfrog" selected through a galax Integrator
    integ = gd.Integrator(dynamics_solver=gd.OrbitSolver(solver=dfx.SemiImplicitEuler(), stepsize_controller=dfx.ConstantStepSize()),
                          diffeq_kw={"max_steps": None, "dt0": 0.1})  # fmt: skip
    orb3 = G["evaluate_orbit"](pot, w0, ts, integrator=integ)
    sie = bd.Integrator(dynamics_solver=bd.OrbitSolver(solver=bd.SemiImplicitEuler(), stepsize_controller=bd.ConstantStepSize()),
                        diffeq_kw={"max_steps": None, "dt0": 0.1})  # fmt: skip
    ref3 = bd.evaluate_orbit(nat, bd.PhaseSpaceCoordinate(q0, p0_kms * u.KMS, 0.0), np.linspace(0.0, 500.0, 6), integrator=sie)
    assert np.array_equal(orb3.q.value, ref3.q)
    with pytest.raises(NotImplementedError, match="dense"):
        G["evaluate_orbit"](pot, w0, ts, dense=True)
    # compute_orbit with a field and with the potential itself
    orb4 = G["compute_orbit"](gd.fields.HamiltonianField(pot), w0, ts)
    ref4 = bd.compute_orbit(nat, bd.PhaseSpaceCoordinate(q0, p0_kms * u.KMS, 0.0), np.linspace(0.0, 500.0, 6))
    assert isinstance(orb4, gd.Orbit) and np.array_equal(orb4.q.value, ref4.q)
    assert np.array_equal(G["compute_orbit"](pot, w0, ts).q.value, ref4.q)

    # the reference's own batch semantics for this call form (one ODE, shared step), on request
    plugin.BATCH = "reference"
    try:
        orb5 = G["evaluate_orbit"](pot, w0, ts)
        ref5 = bd.evaluate_orbit(nat, bd.PhaseSpaceCoordinate(q0, p0_kms * u.KMS, 0.0), np.linspace(0.0, 500.0, 6), joint=True)
        assert np.array_equal(orb5.q.value, ref5.q) and not np.array_equal(orb5.q.value, ref.q)
        assert np.abs(orb5.q.value - ref.q).max() < 1e-4
        assert np.array_equal(G["compute_orbit"](pot, w0, ts).q.value,
                              bd.compute_orbit(nat, bd.PhaseSpaceCoordinate(q0, p0_kms * u.KMS, 0.0), np.linspace(0.0, 500.0, 6), joint=True).q)
        plugin.BATCH = "both"
        with pytest.raises(ValueError):
            G["evaluate_orbit"](pot, w0, ts)
    finally:
        plugin.BATCH = "per-particle"

    # MockStreamGenerator replacement: same call as tests/unit/dynamics/mockstream/test_mockstreamgenerator.py:48-98
    gen = plugin.MockStreamGenerator(gd.FardalStreamDF(), pot)
    prog = gc.PhaseSpaceCoordinate(q=u.Q([30.0, 10, 20], "kpc"), p=u.Q([10.0, -150, -20], "km / s"), t=u.Q(0.0, "Gyr"))
    stream, last = gen.run(np.array([0, 12], dtype=np.uint32), u.Q(np.linspace(0.0, 1000.0, 64), "Myr"), prog, u.Q(1e4, "Msun"))
    assert isinstance(stream, gd.MockStream) and set(stream) == {"lead", "trail"} and isinstance(last, gc.PhaseSpaceCoordinate)
    assert stream["lead"].q.value.shape == (64, 3) and np.isfinite(stream["trail"].p.value).all()
    refgen = bd.MockStreamGenerator(bd.FardalStreamDF(), nat)
    rs, rl = refgen.run(12, np.linspace(0.0, 1000.0, 64), bd.PhaseSpaceCoordinate(np.array([30.0, 10, 20]),
                        np.array([10.0, -150, -20]) * u.KMS, 0.0), 1e4)  # fmt: skip
    assert np.array_equal(stream["lead"].q.value, rs["lead"].q) and np.array_equal(last.q.value, rl.q)
