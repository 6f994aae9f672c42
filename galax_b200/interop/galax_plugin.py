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

This file needs galax, jax, unxt, plum and coordinax at import time; none of them is available in the build
container, so it is exercised only where galax is installed.  ``convert_potential`` (the part that extracts the
parameters from a live galax object) is written against the public attributes cited below.
"""

from __future__ import annotations

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


def register() -> None:
    """Register the plum overloads.  Called on import."""
    import galax.dynamics as gd
    import galax.potential as gp
    import jax
    import jax.numpy as jnp
    from plum import dispatch

    Supported = (
        gp.MilkyWayPotential | gp.MilkyWayPotential2022 | gp.BovyMWPotential2014 | gp.MiyamotoNagaiPotential
        | gp.HernquistPotential | gp.NFWPotential | gp.PowerLawCutoffPotential | gp.MN3Sech2Potential
        | gp.MN3ExponentialPotential | gp.LM10Potential | gp.KeplerPotential | gp.PlummerPotential | gp.KuzminPotential
        | gp.IsochronePotential | gp.SatohPotential | gp.JaffePotential | gp.BurkertPotential
        | gp.StoneOstriker15Potential | gp.TriaxialHernquistPotential | gp.LMJ09LogarithmicPotential
        | gp.LogarithmicPotential
    )  # fmt: skip

    def _np(x):
        if isinstance(x, jax.core.Tracer):
            raise NotImplementedError(
                "galax_b200 kernels cannot be traced/differentiated by JAX; call them outside jit/grad "
                "or use the XLA-FFI build (INTEGRATION.md section 3)"
            )
        return np.asarray(x, dtype=np.float64)

    # potential/_src/register_funcs.py:86-98 (array, array) forms
    @dispatch
    def gradient(pot: Supported, xyz: jax.Array | np.ndarray, t: object, /):
        return jnp.asarray(convert_potential(pot).gradient(_np(xyz), t))

    @dispatch
    def hessian(pot: Supported, xyz: jax.Array | np.ndarray, t: object, /):
        return jnp.asarray(convert_potential(pot).hessian(_np(xyz), t))

    @dispatch
    def acceleration(pot: Supported, xyz: jax.Array | np.ndarray, t: object, /):
        return jnp.asarray(convert_potential(pot).acceleration(_np(xyz), t))

    # galax registers every overload on plum's global dispatcher, keyed by the function NAME
    # (potential/_src/register_funcs.py:11,33,86,...; interop/astropy/dynamics.py:8,21): defining functions
    # called ``gradient`` / ``hessian`` / ``acceleration`` / ``evaluate_orbit`` under ``@dispatch`` here adds
    # methods to the same generics, and the narrower ``Supported`` annotation wins by specificity.

    # dynamics/_src/legacy/funcs.py:42-51 -- evaluate_orbit(pot, w0, t, *, integrator=None, dense=False)
    def _solver_spec(integrator):
        import diffrax as dfx

        if integrator is None:
            return bd.Integrator()
        ds = integrator.dynamics_solver
        ctrl = ds.stepsize_controller
        if isinstance(ds.solver, dfx.Dopri8) and isinstance(ctrl, dfx.PIDController):
            c = bd.PIDController(rtol=float(ctrl.rtol), atol=float(ctrl.atol), pcoeff=float(ctrl.pcoeff),
                                 icoeff=float(ctrl.icoeff), dcoeff=float(ctrl.dcoeff), dtmin=ctrl.dtmin,
                                 dtmax=ctrl.dtmax, force_dtmin=bool(ctrl.force_dtmin),
                                 factormin=float(ctrl.factormin), factormax=float(ctrl.factormax),
                                 safety=float(ctrl.safety))  # fmt: skip
            s = bd.Dopri8()
        elif isinstance(ds.solver, dfx.SemiImplicitEuler) and isinstance(ctrl, dfx.ConstantStepSize):
            c, s = bd.ConstantStepSize(), bd.SemiImplicitEuler()
        elif isinstance(ds.solver, dfx.LeapfrogMidpoint) and isinstance(ctrl, dfx.ConstantStepSize):
            c, s = bd.ConstantStepSize(), bd.LeapfrogMidpoint()
        else:
            raise NotImplementedError(f"{type(ds.solver).__name__} / {type(ctrl).__name__} has no galax_b200 kernel")
        return bd.Integrator(dynamics_solver=bd.OrbitSolver(solver=s, stepsize_controller=c, max_steps=ds.max_steps),
                             diffeq_kw=dict(integrator.diffeq_kw))  # fmt: skip

    @dispatch  # same mechanism as galax/interop/astropy/dynamics.py:21-93
    def evaluate_orbit(pot: Supported, w0: object, t: object, /, *, integrator: object = None, dense: bool = False):
        import galax.coordinates as gc
        import unxt as u

        if dense:
            raise NotImplementedError("dense=True is not supported by galax_b200")
        units = pot.units
        tt = _np(u.ustrip(units["time"], t))
        if isinstance(w0, gc.AbstractPhaseSpaceObject):
            q, p = w0._qp(units=units)  # coordinates/_src/base.py:348-386
            w = bd.PhaseSpaceCoordinate(_np(q), _np(p), None if getattr(w0, "t", None) is None
                                        else _np(u.ustrip(units["time"], w0.t)))  # fmt: skip
        else:
            w = w0 if isinstance(w0, tuple) else _np(w0)
        orb = bd.evaluate_orbit(convert_potential(pot), w, tt, integrator=_solver_spec(integrator))
        # re-wrap exactly as legacy/funcs.py:31-39 (orbit_from_psp) / orbit/register_dfx.py:91-96
        return gd.Orbit(q=u.Q(jnp.asarray(orb.q), units["length"]), p=u.Q(jnp.asarray(orb.p), units["speed"]),
                        t=u.Q(jnp.asarray(orb.t), units["time"]), frame=gc.frames.simulation_frame)  # fmt: skip


try:  # pragma: no cover - needs galax
    register()
except ImportError as e:  # galax (or jax / unxt / plum) not installed: the plugin is inert
    _IMPORT_ERROR = e
