"""galax plugin: route the supported potentials / solvers to libgalax_b200.so without editing galax.

How galax finds it (``/root/reference/src/galax/potential/setup_package.py:61-84``): every portion of galax ends
its ``__init__`` with ``load_interop_plugins("galax.<portion>.interop")``, which imports each module registered
under that entry-point group.  A distribution shipping this module declares

    [project.entry-points."galax.potential.interop"]
    galax_b200 = "galax_b200.interop.galax_plugin"
    [project.entry-points."galax.dynamics.interop"]
    galax_b200 = "galax_b200.interop.galax_plugin"

Importing it registers more-specific plum dispatches (exactly what ``galax/interop/astropy/dynamics.py:21-93``
does in-tree), which win over the generic ones by type specificity.

This file needs galax, jax, unxt, plum, coordinax and diffrax at import time.  None of them can be installed in the
build container, so ``tests/test_interop_plugin.py`` imports it against duck-typed stand-ins (``tests/fake_galax``:
same public attribute names and call forms, cited per class) and checks that ``convert_potential`` produces the same
``gx_potential`` bytes as the native constructors and that every overload below is registered, dispatches, and
returns the reference's container types.  Everything here touches galax objects only through the public attributes
cited next to each use.

Overloads registered (reference signature -> here):
  gradient / hessian        potential/_src/register_funcs.py:86-155, 276-320   all six / four call forms
  acceleration              :327-340 (array form registered for speed; the reference's wrapper `-gradient` covers the rest)
  tidal_tensor              :347-377 is a wrapper over `api.hessian`: it lands on the hessian overloads by itself
  evaluate_orbit            dynamics/_src/legacy/funcs.py:42-254
  compute_orbit             dynamics/_src/orbit/compute.py:28-98 (both w0 kinds)
  MockStreamGenerator       a same-signature replacement class (mockstream_generator.py:33-275): its inner solves run
                            under jax.vmap / lax.scan in the reference, which an eager binding cannot intercept
"""

# Batch semantics of the adaptive solves (GALAX_B200_BATCH, or `galax_plugin.BATCH = ...` at run time):
#   "per-particle" (default)  every orbit controls its own step -- what north_star asks of the kernels and what the
#                             reference does under vmap / lstrat.VMap / batched start times;
#   "reference"               a batch of initial conditions with scalar times is ONE ODE with a shared adaptive step and
#                             an error norm over all 6N components, exactly as the reference's evaluate_orbit /
#                             compute_orbit do for that call form (orbit/solver.py:774-803, legacy/integrator.py:288-298):
#                             gx_integrate_adaptive_joint.  Use it when the reference's numbers for that call form are
#                             wanted digit for digit (tests/test_gpu_joint.py reproduces its 8-digit doctests).
# The two agree to the tolerance of the JOINT norm, which lets single orbits of a large batch err by more than rtol.
import os

BATCH = os.environ.get("GALAX_B200_BATCH", "per-particle")


def _joint() -> bool:
    if BATCH not in ("per-particle", "reference"):
        raise ValueError(f"GALAX_B200_BATCH / galax_plugin.BATCH must be 'per-particle' or 'reference', not {BATCH!r}")
    return BATCH == "reference"


# NOTE: no `from __future__ import annotations` here.  The overloads below are annotated with types that only exist
# inside register() (the `Supported` union of galax classes); plum resolves annotations when a method is registered,
# and a postponed (string) annotation naming a local would be unresolvable.

import numpy as np

from .. import dynamics as bd
from .. import potential as bp


def _value(param, unit, time_unit=None):
    """``ConstantParameter`` -> float in ``unit``; ``LinearParameter`` -> ``galax_b200.potential.LinearParameter`` with
    slope in ``unit``/``time_unit`` (the integrators evaluate it per stage).  Anything else (``UserParameter``, ...) is a
    general function of time and unsupported.

    galax: ``potential/_src/params/constant.py`` (``ConstantParameter.value``), ``params/core.py:25-110``
    (``LinearParameter.slope / point_time / point_value``), ``params/field.py:206-223``.
    """
    import unxt as u

    name = type(param).__name__
    if name == "ConstantParameter":
        return float(u.ustrip(unit, param.value))
    if name == "LinearParameter" and time_unit is not None:
        return bp.LinearParameter(slope=float(u.ustrip(unit / time_unit, param.slope)),
                                  point_time=float(u.ustrip(time_unit, param.point_time)),
                                  point_value=float(u.ustrip(unit, param.point_value)))
    raise NotImplementedError(f"galax_b200 supports ConstantParameter and LinearParameter only, got {name}")


def convert_potential(pot) -> bp.AbstractPotential:
    """galax potential object -> galax_b200 potential with the same parameters and the same ``G``.

    Reads ``pot.constants["G"].value`` (never recomputed; ``potential/_src/base.py:35,91-95``) and each
    parameter in the potential's own unit system (``pot.units``), as the reference's ``_potential`` methods do
    (e.g. ``builtin/miyamotonagai.py:59-66``).
    """
    import galax.potential as gp

    G = float(pot.constants["G"].value)
    ul, um, ut = pot.units["length"], pot.units["mass"], pot.units["time"]
    if isinstance(pot, gp.MiyamotoNagaiPotential):
        return bp.MiyamotoNagaiPotential(_value(pot.m_tot, um, ut), _value(pot.a, ul, ut), _value(pot.b, ul, ut), G=G)
    if isinstance(pot, gp.HernquistPotential):
        return bp.HernquistPotential(_value(pot.m_tot, um, ut), _value(pot.r_s, ul, ut), G=G)
    if isinstance(pot, gp.KeplerPotential):
        return bp.KeplerPotential(_value(pot.m_tot, um, ut), G=G)
    if isinstance(pot, gp.NFWPotential):
        return bp.NFWPotential(_value(pot.m, um, ut), _value(pot.r_s, ul, ut), G=G)
    if isinstance(pot, gp.PowerLawCutoffPotential):
        return bp.PowerLawCutoffPotential(
            _value(pot.m_tot, um), _value(pot.alpha, pot.units["dimensionless"]), _value(pot.r_c, ul), G=G
        )
    if isinstance(pot, (gp.MN3Sech2Potential, gp.MN3ExponentialPotential)):
        cls = bp.MN3Sech2Potential if isinstance(pot, gp.MN3Sech2Potential) else bp.MN3ExponentialPotential
        return cls(_value(pot.m_tot, um), _value(pot.h_R, ul), _value(pot.h_z, ul),
                   positive_density=bool(pot.positive_density), G=G)  # fmt: skip
    # further single components: (galax class name, galax_b200 class, ((field, unit key), ...)) -- positional order of
    # the galax_b200 constructors; units as the reference's ``_potential`` methods strip them
    ua, ud, us = pot.units["angle"], pot.units["dimensionless"], pot.units["speed"]
    simple = (
        ("PlummerPotential", bp.PlummerPotential, (("m_tot", um), ("r_s", ul))),
        ("KuzminPotential", bp.KuzminPotential, (("m_tot", um), ("r_s", ul))),
        ("IsochronePotential", bp.IsochronePotential, (("m_tot", um), ("r_s", ul))),
        ("SatohPotential", bp.SatohPotential, (("m_tot", um), ("a", ul), ("b", ul))),
        ("JaffePotential", bp.JaffePotential, (("m_tot", um), ("r_s", ul))),
        ("BurkertPotential", bp.BurkertPotential, (("m", um), ("r_s", ul))),
        ("StoneOstriker15Potential", bp.StoneOstriker15Potential, (("m_tot", um), ("r_c", ul), ("r_h", ul))),
        ("TriaxialHernquistPotential", bp.TriaxialHernquistPotential, (("m_tot", um), ("r_s", ul), ("q1", ud), ("q2", ud))),
        ("LMJ09LogarithmicPotential", bp.LMJ09LogarithmicPotential,
         (("v_c", us), ("r_s", ul), ("q1", ud), ("q2", ud), ("q3", ud), ("phi", ua))),
        ("LogarithmicPotential", bp.LogarithmicPotential, (("v_c", us), ("r_s", ul))),
    )  # fmt: skip
    for name, cls, fields in simple:
        if type(pot).__name__ == name:
            return cls(*[_value(getattr(pot, f), unit, ut) for f, unit in fields], G=G)
    if type(pot).__name__ == "NullPotential":
        return bp.NullPotential(G=G)
    if isinstance(pot, gp.AbstractCompositePotential):  # CompositePotential and the pre-composited MW models
        return bp.CompositePotential({k: convert_potential(v) for k, v in pot.items()}, G=G)
    raise NotImplementedError(f"{type(pot).__name__} has no galax_b200 kernel")


def _xyz_t(pot, q, t):
    """(xyz [*batch, 3], t) as plain fp64 arrays in the potential's units: what ``parse_to_xyz_t(None, q, t,
    ustrip=pot.units, dtype=float)`` returns (potential/_src/utils.py:308-396).  The real parser is used when galax
    provides it; the fallback handles arrays, Quantities, Cartesian position vectors and phase-space objects."""
    import unxt as u

    try:
        from galax.potential._src.utils import parse_to_xyz_t  # type: ignore[import-not-found]

        xyz, tt = parse_to_xyz_t(None, q, t, ustrip=pot.units, dtype=float)
        return np.asarray(xyz, dtype=np.float64), (0.0 if tt is None else float(np.asarray(tt).reshape(-1)[0]))
    except ImportError:
        pass
    if t is None and hasattr(q, "t"):
        t = q.t
    if hasattr(q, "q") and hasattr(q, "p"):  # phase-space object: its position
        q = q.q
    if hasattr(q, "xyz"):  # Cartesian position vector
        q = q.xyz
    xyz = u.ustrip(pot.units["length"], q) if isinstance(q, u.AbstractQuantity) else np.asarray(q, dtype=np.float64)
    tt = 0.0 if t is None else (u.ustrip(pot.units["time"], t) if isinstance(t, u.AbstractQuantity) else t)
    return np.asarray(xyz, dtype=np.float64), float(np.asarray(tt, dtype=np.float64).reshape(-1)[0])


def _integrator_spec(integrator):
    """galax ``Integrator`` (legacy/integrator.py:42-245) -> galax_b200 ``Integrator`` with the same solver, controller
    coefficients, ``max_steps`` and ``diffeq_kw``; unsupported solver / controller pairs raise."""
    if integrator is None:
        return bd.Integrator()
    return bd.Integrator(dynamics_solver=_solver_spec(integrator.dynamics_solver), diffeq_kw=dict(integrator.diffeq_kw))


def _solver_spec(ds):
    """galax ``OrbitSolver`` (orbit/solver.py:121-141) -> galax_b200 ``OrbitSolver``."""
    import diffrax as dfx

    if ds is None:
        return bd.OrbitSolver()
    ctrl = ds.stepsize_controller
    if isinstance(ds.solver, (dfx.Dopri8, dfx.Dopri5)) and isinstance(ctrl, dfx.PIDController):
        c = bd.PIDController(rtol=float(ctrl.rtol), atol=float(ctrl.atol), pcoeff=float(ctrl.pcoeff),
                             icoeff=float(ctrl.icoeff), dcoeff=float(ctrl.dcoeff), dtmin=ctrl.dtmin,
                             dtmax=ctrl.dtmax, force_dtmin=bool(ctrl.force_dtmin),
                             factormin=float(ctrl.factormin), factormax=float(ctrl.factormax),
                             safety=float(ctrl.safety))  # fmt: skip
        s = bd.Dopri8() if isinstance(ds.solver, dfx.Dopri8) else bd.Dopri5()
    elif isinstance(ds.solver, dfx.SemiImplicitEuler) and isinstance(ctrl, dfx.ConstantStepSize):
        c, s = bd.ConstantStepSize(), bd.SemiImplicitEuler()
    elif isinstance(ds.solver, dfx.LeapfrogMidpoint) and isinstance(ctrl, dfx.ConstantStepSize):
        c, s = bd.ConstantStepSize(), bd.LeapfrogMidpoint()
    else:
        raise NotImplementedError(f"{type(ds.solver).__name__} / {type(ctrl).__name__} has no galax_b200 kernel")
    return bd.OrbitSolver(solver=s, stepsize_controller=c, max_steps=ds.max_steps)


def _w0(w0, units):
    """galax initial conditions (legacy/funcs.py:45: phase-space object, (q, p) tuple, (*batch, 6) array) -> what
    ``galax_b200.dynamics`` takes, in the potential's units; also returns the frame to re-attach."""
    import galax.coordinates as gc
    import unxt as u

    if isinstance(w0, gc.AbstractPhaseSpaceObject):
        q, p = w0._qp(units=units)  # coordinates/_src/base.py:348-386
        t = getattr(w0, "t", None)
        return bd.PhaseSpaceCoordinate(np.asarray(u.ustrip(units["length"], q), dtype=np.float64),
                                       np.asarray(u.ustrip(units["speed"], p), dtype=np.float64),
                                       None if t is None else float(np.asarray(u.ustrip(units["time"], t)))), w0.frame  # fmt: skip
    frame = gc.frames.simulation_frame
    if isinstance(w0, tuple):
        return (np.asarray(w0[0], dtype=np.float64), np.asarray(w0[1], dtype=np.float64)), frame
    return np.asarray(w0, dtype=np.float64), frame


def _wrap_orbit(orb, units, frame):
    """Raw arrays -> ``gd.Orbit`` exactly as legacy/funcs.py:31-39 (orbit_from_psp) / orbit/register_dfx.py:91-96."""
    import galax.dynamics as gd
    import jax.numpy as jnp
    import unxt as u

    return gd.Orbit(q=u.Q(jnp.asarray(orb.q), units["length"]), p=u.Q(jnp.asarray(orb.p), units["speed"]),
                    t=u.Q(jnp.asarray(orb.t), units["time"]), frame=frame)  # fmt: skip


class MockStreamGenerator:
    """Same constructor and ``run`` signature as ``gd.MockStreamGenerator`` (mockstream_generator.py:33-60,160-169),
    backed by the device pipeline of ``galax_b200.dynamics.MockStreamGenerator`` (progenitor orbit -> release kernel ->
    one work-queue launch for both arms).  ``rng`` is a jax PRNG key: its raw key data seed the threefry restatement,
    so the stream is the one the reference draws (``FardalStreamDF``)."""

    def __init__(self, df, potential, *, progenitor_integrator=None, stream_integrator=None):
        name = type(df).__name__
        if name not in ("FardalStreamDF", "ChenStreamDF"):
            raise NotImplementedError(f"{name} has no galax_b200 release kernel")
        self.df, self.potential = df, potential
        self._gen = bd.MockStreamGenerator(
            bd.FardalStreamDF() if name == "FardalStreamDF" else bd.ChenStreamDF(), convert_potential(potential),
            progenitor_integrator=_integrator_spec(progenitor_integrator),
            stream_integrator=_integrator_spec(stream_integrator))  # fmt: skip

    @property
    def units(self):
        return self.potential.units

    def run(self, rng, ts, prog_w0, prog_mass, *, vmapped=None):
        import galax.coordinates as gc
        import galax.dynamics as gd
        import jax.numpy as jnp
        import unxt as u

        units = self.units
        try:  # a typed jax key -> its uint32[2] data (jax.random.key_data); raw key data / seeds pass through
            import jax.random as jr

            rng = np.asarray(jr.key_data(rng), dtype=np.uint32)
        except (ImportError, AttributeError, TypeError):
            pass
        tsn = np.asarray(u.ustrip(units["time"], ts), dtype=np.float64)
        w0, frame = _w0(prog_w0, units)
        if callable(prog_mass):
            raise NotImplementedError("a time-dependent progenitor mass (ProgenitorMassCallable) is not supported")
        mass = float(np.asarray(u.ustrip(units["mass"], prog_mass)))
        stream, prog = self._gen.run(rng, tsn, w0, mass, vmapped=vmapped)
        comps = {}
        for name, arm in stream.items():
            comps[name] = gd.MockStreamArm(q=u.Q(jnp.asarray(arm.q), units["length"]), p=u.Q(jnp.asarray(arm.p), units["speed"]),
                                           t=u.Q(jnp.asarray(arm.t), units["time"]),
                                           release_time=u.Q(jnp.asarray(arm.release_time), units["time"]), frame=frame)  # fmt: skip
        last = gc.PhaseSpaceCoordinate(q=u.Q(jnp.asarray(prog.q), units["length"]), p=u.Q(jnp.asarray(prog.p), units["speed"]),
                                       t=u.Q(jnp.asarray(prog.t), units["time"]), frame=frame)  # fmt: skip
        return gd.MockStream(comps), last


def register() -> None:
    """Register the plum overloads.  Called on import."""
    import coordinax as cx
    import galax.coordinates as gc
    import galax.dynamics as gd
    import galax.potential as gp
    import jax
    import jax.numpy as jnp
    import unxt as u
    from plum import dispatch

    Supported = (
        gp.MilkyWayPotential | gp.MilkyWayPotential2022 | gp.BovyMWPotential2014 | gp.MiyamotoNagaiPotential
        | gp.HernquistPotential | gp.NFWPotential | gp.PowerLawCutoffPotential | gp.MN3Sech2Potential
        | gp.MN3ExponentialPotential | gp.LM10Potential | gp.KeplerPotential | gp.PlummerPotential | gp.KuzminPotential
        | gp.IsochronePotential | gp.SatohPotential | gp.JaffePotential | gp.BurkertPotential
        | gp.StoneOstriker15Potential | gp.TriaxialHernquistPotential | gp.LMJ09LogarithmicPotential
        | gp.LogarithmicPotential | gp.CompositePotential
    )  # fmt: skip
    ArrayLike = jax.Array | np.ndarray | list | tuple | float | int

    def _np(x):
        if isinstance(x, jax.core.Tracer):
            raise NotImplementedError(
                "galax_b200 kernels cannot be traced/differentiated by JAX; call them outside jit/grad "
                "or use the XLA-FFI build (INTEGRATION.md section 3)"
            )
        return np.asarray(x, dtype=np.float64)

    def _t(t):
        return float(np.asarray(_np(t)).reshape(-1)[0])

    # galax registers every overload on plum's global dispatcher, keyed by the function NAME
    # (potential/_src/register_funcs.py:11,33,86,...; interop/astropy/dynamics.py:8,21): defining functions called
    # ``gradient`` / ``hessian`` / ... under ``@dispatch`` here adds methods to the same generics, and the narrower
    # ``Supported`` annotation wins by specificity.

    # ---- gradient: potential/_src/register_funcs.py:86-155 ----
    @dispatch
    def gradient(pot: Supported, xyz: ArrayLike, t: ArrayLike, /):  # :86-98 arrays in, bare array out
        return jnp.asarray(convert_potential(pot).gradient(_np(xyz), _t(t)))

    @dispatch
    def gradient(pot: Supported, xyz: ArrayLike, /, *, t: ArrayLike):  # noqa: F811  :102-107
        return jnp.asarray(convert_potential(pot).gradient(_np(xyz), _t(t)))

    @dispatch
    def gradient(pot: Supported, xyz: u.AbstractQuantity, /, *, t: u.AbstractQuantity):  # noqa: F811  :114-123
        x, tt = _xyz_t(pot, xyz, t)
        return u.Q.from_(jnp.asarray(convert_potential(pot).gradient(x, tt)), pot.units["acceleration"])

    @dispatch
    def gradient(pot: Supported, q: u.AbstractQuantity, t: u.AbstractQuantity, /):  # noqa: F811  :126-134
        x, tt = _xyz_t(pot, q, t)
        return u.Q.from_(jnp.asarray(convert_potential(pot).gradient(x, tt)), pot.units["acceleration"])

    @dispatch
    def gradient(pot: Supported, tq: object, /, *, t: object = None):  # noqa: F811  :140-148 vectors / phase-space objects
        x, tt = _xyz_t(pot, tq, t)
        return cx.vecs.CartesianAcc3D.from_(jnp.asarray(convert_potential(pot).gradient(x, tt)), pot.units["acceleration"])

    @dispatch
    def gradient(pot: Supported, q: object, t: object, /):  # noqa: F811  :151-155
        x, tt = _xyz_t(pot, q, t)
        return cx.vecs.CartesianAcc3D.from_(jnp.asarray(convert_potential(pot).gradient(x, tt)), pot.units["acceleration"])

    # ---- hessian: :276-320 ----
    @dispatch
    def hessian(pot: Supported, xyz: ArrayLike, t: ArrayLike, /):  # :276-288
        return jnp.asarray(convert_potential(pot).hessian(_np(xyz), _t(t)))

    @dispatch
    def hessian(pot: Supported, xyz: ArrayLike, /, *, t: ArrayLike):  # noqa: F811  :292-297
        return jnp.asarray(convert_potential(pot).hessian(_np(xyz), _t(t)))

    @dispatch
    def hessian(pot: Supported, tq: object, /, *, t: object = None):  # noqa: F811  :303-310 -> Quantity["frequency drift"]
        x, tt = _xyz_t(pot, tq, t)
        return u.Q(jnp.asarray(convert_potential(pot).hessian(x, tt)), pot.units["frequency drift"])

    @dispatch
    def hessian(pot: Supported, q: object, t: object, /):  # noqa: F811  :313-320
        x, tt = _xyz_t(pot, q, t)
        return u.Q(jnp.asarray(convert_potential(pot).hessian(x, tt)), pot.units["frequency drift"])

    # ---- acceleration (:327-340 is `-gradient(...)`; the array form directly, one kernel pass, no negation on the host)
    @dispatch
    def acceleration(pot: Supported, xyz: ArrayLike, t: ArrayLike, /):
        return jnp.asarray(convert_potential(pot).acceleration(_np(xyz), _t(t)))

    # ---- tidal_tensor (:347-377): hessian minus a third of its trace; the array form in one pass
    @dispatch
    def tidal_tensor(pot: Supported, xyz: ArrayLike, t: ArrayLike, /):
        return jnp.asarray(convert_potential(pot).tidal_tensor(_np(xyz), _t(t)))

    # ---- evaluate_orbit: dynamics/_src/legacy/funcs.py:42-51, and the t= keyword form :216-254 ----
    def _evaluate_orbit(pot, w0, t, integrator, dense):
        if dense:
            raise NotImplementedError("dense=True (an interpolated Orbit) is not supported by galax_b200")
        units = pot.units
        tt = np.atleast_1d(_np(u.ustrip(units["time"], t)))
        w, frame = _w0(w0, units)
        orb = bd.evaluate_orbit(convert_potential(pot), w, tt, integrator=_integrator_spec(integrator), joint=_joint())
        return _wrap_orbit(orb, units, frame)

    @dispatch  # same mechanism as galax/interop/astropy/dynamics.py:21-93
    def evaluate_orbit(pot: Supported, w0: object, t: object, /, *, integrator: object = None, dense: bool = False):
        return _evaluate_orbit(pot, w0, t, integrator, dense)

    @dispatch
    def evaluate_orbit(pot: Supported, w0: object, /, *, t: object, integrator: object = None, dense: bool = False):  # noqa: F811
        return _evaluate_orbit(pot, w0, t, integrator, dense)

    # ---- compute_orbit: dynamics/_src/orbit/compute.py:28-98 ----
    # (batch semantics: see BATCH at the top of this module)
    @dispatch
    def compute_orbit(field: Supported | gd.fields.HamiltonianField, w0: gc.AbstractPhaseSpaceObject, ts: object, /, *,
                      solver: object = None, dense: bool = False):
        if dense:
            raise NotImplementedError("dense=True (an interpolated Orbit) is not supported by galax_b200")
        pot = field.potential if isinstance(field, gd.fields.HamiltonianField) else field
        if not isinstance(pot, Supported):
            raise NotImplementedError(f"{type(pot).__name__} has no galax_b200 kernel")
        units = pot.units
        tt = np.atleast_1d(_np(u.ustrip(units["time"], ts)))
        w, frame = _w0(w0, units)
        orb = bd.compute_orbit(convert_potential(pot), w, tt, solver=_solver_spec(solver), joint=_joint())
        return _wrap_orbit(orb, units, frame)


try:  # pragma: no cover - needs galax
    register()
except ImportError as e:  # galax (or jax / unxt / plum) not installed: the plugin is inert
    _IMPORT_ERROR = e
