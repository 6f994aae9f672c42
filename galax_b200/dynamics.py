"""Host-side mirror of ``galax.dynamics`` for the hot path: orbit integration and mock streams.

Names, argument order and defaults follow the reference:

* ``evaluate_orbit(pot, w0, t, *, integrator=None, dense=False)``  legacy/funcs.py:42-213
* ``compute_orbit(pot_or_field, w0, ts, *, solver=None, dense=False)``  orbit/compute.py:28-98
* ``Integrator(dynamics_solver=..., diffeq_kw=...)``  legacy/integrator.py:42-245 (default Dopri8, rtol=atol=1e-7,
  ``max_steps=None``)
* ``OrbitSolver(solver, stepsize_controller, max_steps)``  orbit/solver.py:121-141 (default Dopri8, 1e-8, 2**16)
* ``MockStreamGenerator(df, potential, progenitor_integrator=, stream_integrator=).run(...)``
  legacy/mockstream/mockstream_generator.py:160-275, ``FardalStreamDF`` / ``ChenStreamDF`` df/*.py

``SemiImplicitEuler``, ``LeapfrogMidpoint``, ``Dopri8``, ``ConstantStepSize`` and ``PIDController`` are
parameter records standing in for the diffrax objects the reference passes around (diffrax is not a
dependency here); their fields and defaults are diffrax 0.7.0's.

All numerics run in the CUDA library (``include/galax_b200.h``); arrays are fp64 in the potential's unit
system.  numpy in -> numpy out; CUDA torch tensors in -> CUDA tensors out with no host round trip.

Semantic notes (also in DESIGN.md):
* step control is always per particle (what the reference does on its ``vmap`` / vectorised paths); the
  reference's scalar-time batch path instead solves the whole batch as one ODE with a shared step.
* ``dense=True``, events, ``args`` and differentiating through the call are not supported and raise.
"""

from __future__ import annotations

import ctypes as C
import dataclasses
import math
from typing import Any, Callable, Sequence

import numpy as np

from . import _lib
from .potential import AbstractPotential, _to_device, _to_host_many

# ------------------------------------------------------------------------------------------------
# diffrax stand-ins (plain records)


@dataclasses.dataclass(frozen=True)
class SemiImplicitEuler:
    """diffrax.SemiImplicitEuler: q1 = q0 + p0 h ; p1 = p0 - grad Phi(q1) h.  The reference's "leapfrog".

    ``strict=True`` (not a diffrax field) selects the reference-order kernel (``GX_SCHEME_STRICT``): the reference's
    arithmetic operation by operation -- components summed in composite order, IEEE division / square root, no FMA
    contraction, portable log1p -- reproducible bit for bit on a CPU, several times slower."""

    strict: bool = False


@dataclasses.dataclass(frozen=True)
class LeapfrogMidpoint:
    """diffrax.LeapfrogMidpoint: y_{n+1} = y_{n-1} + f(y_n) (t_{n+1} - t_{n-1}).  ``strict``: see SemiImplicitEuler."""

    strict: bool = False


@dataclasses.dataclass(frozen=True)
class Dopri8:
    """diffrax.Dopri8 (Prince-Dormand 8(7)13M + FSAL); ``scan_kind`` has no numerical effect.

    ``strict=True`` (not a diffrax field): the reference-order kernel (``GX_SOLVER_STRICT``) -- generic Runge-Kutta form,
    IEEE division / square root, no FMA contraction, portable transcendental functions: results and accept / reject
    sequence reproducible bit for bit on a CPU; roughly an order of magnitude slower than the default kernel."""

    scan_kind: str | None = None
    strict: bool = False


@dataclasses.dataclass(frozen=True)
class Dopri5:
    """diffrax.Dopri5 (Dormand-Prince 5(4) + FSAL): default of the reference's experimental StreamSimulator.
    ``strict``: see Dopri8."""

    scan_kind: str | None = None
    strict: bool = False


@dataclasses.dataclass(frozen=True)
class ConstantStepSize:
    """diffrax.ConstantStepSize: needs ``dt0``."""


@dataclasses.dataclass(frozen=True)
class PIDController:
    """diffrax.PIDController with diffrax 0.7.0 defaults."""

    rtol: float
    atol: float
    pcoeff: float = 0.0
    icoeff: float = 1.0
    dcoeff: float = 0.0
    dtmin: float | None = None
    dtmax: float | None = None
    force_dtmin: bool = True
    factormin: float = 0.2
    factormax: float = 10.0
    safety: float = 0.9

    def c_struct(self, dt0: float | None) -> _lib.GxPid:
        if self.dtmin is not None and not self.force_dtmin:
            raise NotImplementedError("PIDController(force_dtmin=False) is not supported")
        pid = _lib.GxPid()
        pid.rtol, pid.atol = float(self.rtol), float(self.atol)
        pid.pcoeff, pid.icoeff, pid.dcoeff = float(self.pcoeff), float(self.icoeff), float(self.dcoeff)
        pid.safety, pid.factormin, pid.factormax = float(self.safety), float(self.factormin), float(self.factormax)
        pid.dtmin = -1.0 if self.dtmin is None else float(self.dtmin)
        pid.dtmax = -1.0 if self.dtmax is None else float(self.dtmax)
        pid.force_dtmin = int(self.force_dtmin)
        pid.dt0 = -1.0 if dt0 is None else float(dt0)
        return pid


# ------------------------------------------------------------------------------------------------
# containers (coordinates/_src/pscs, dynamics/_src/orbit/orbit.py:18-75)


class _PhaseSpaceDiagnostics:
    """coordinates/_src/pscs/base.py:182-330: dynamical quantities of (q, p), evaluated on the device in one pass
    (``gx_energy_angmom``) for any leading shape."""

    def kinetic_energy(self):
        """|p|^2 / 2 (pscs/base.py, ``kinetic_energy``)."""
        return _energy(None, self.q, self.p)

    def potential_energy(self, potential=None):
        """Phi(q) (pscs/base.py:182-229)."""
        pot = _as_potential(potential if potential is not None else getattr(self, "potential", None))
        return pot.potential(self.q, 0.0)

    def _fused(self, name, potential=None):
        """A diagnostic the integrator kernel already evaluated at every save (``diagnostics=`` of evaluate_orbit /
        compute_orbit / OrbitSolver.solve), if it is the one asked for."""
        d = getattr(self, "diagnostics", None) or {}
        if name in d and (potential is None or potential is getattr(self, "potential", None)):
            return d[name]
        return None

    def total_energy(self, potential=None):
        """|p|^2/2 + Phi(q) (pscs/base.py:231-283)."""
        fused = self._fused("energy", potential)
        if fused is not None:
            return fused
        pot = _as_potential(potential if potential is not None else getattr(self, "potential", None))
        return _energy(pot, self.q, self.p, t=getattr(self, "t", None) if pot.is_time_dependent else None)

    def angular_momentum(self):
        """q x p (pscs/base.py:285-322, ``specific_angular_momentum``)."""
        fused = self._fused("angular_momentum")
        return fused if fused is not None else _energy(None, self.q, self.p, want="L")

    def tidal_tensor(self, potential=None):
        """J - tr(J)/3 I of the potential's Hessian along the states (potential/_src/register_funcs.py:347-377)."""
        fused = self._fused("tidal_tensor", potential)
        if fused is not None:
            return fused
        pot = _as_potential(potential if potential is not None else getattr(self, "potential", None))
        return pot.tidal_tensor(self.q, 0.0)

    def w(self):
        return _cat(self.q, self.p)


@dataclasses.dataclass
class PhaseSpacePosition(_PhaseSpaceDiagnostics):
    q: Any
    p: Any


@dataclasses.dataclass
class PhaseSpaceCoordinate(_PhaseSpaceDiagnostics):
    q: Any
    p: Any
    t: Any = None


@dataclasses.dataclass
class Orbit(_PhaseSpaceDiagnostics):
    """q, p: (*batch, T, 3); t: (T,).  Mirrors ``gd.Orbit`` (orbit/orbit.py:18-75)."""

    q: Any
    p: Any
    t: Any
    potential: AbstractPotential | None = None
    status: Any = None
    n_steps: Any = None
    diagnostics: dict | None = None  # fused kernel epilogue: {"energy": (*batch, T), "angular_momentum": ..., "tidal_tensor": ...}

    @property
    def shape(self):
        return tuple(self.q.shape[:-1])

    def __getitem__(self, idx):
        return PhaseSpaceCoordinate(self.q[..., idx, :], self.p[..., idx, :], self.t[idx])


@dataclasses.dataclass
class Solution:
    """Shape-compatible stand-in for ``diffrax.Solution``: ys = (q, p) each (*batch, T, 3)."""

    t0: Any
    t1: Any
    ts: Any
    ys: tuple
    stats: dict
    result: Any


def _cat0(a, b):
    if isinstance(a, np.ndarray):
        return np.concatenate([a, b], axis=0)
    import torch

    return torch.cat([a, b], dim=0)


def _cat(q, p):
    if isinstance(q, np.ndarray):
        return np.concatenate([q, p], axis=-1)
    import torch

    return torch.cat([q, p], dim=-1)


class HamiltonianField:
    """``gd.fields.HamiltonianField(pot)`` (orbit/field_hamiltonian.py:229-249): (dq, dp) = (p, -grad Phi)."""

    def __init__(self, potential: AbstractPotential):
        if not isinstance(potential, AbstractPotential):
            raise TypeError("HamiltonianField needs a galax_b200 potential")
        self.potential = potential

    def __call__(self, t, *xv):
        """The reference's array-level call forms (field_hamiltonian.py:108-175): ``field(t, q, p[, None])``,
        ``field(t, (q, p)[, None])`` and ``field(t, qp[, None])`` with ``qp`` of shape ``(*batch, 6)``."""
        if xv and xv[-1] is None and len(xv) > 1:
            xv = xv[:-1]  # args=None
        if len(xv) == 2:
            q, p = xv
        elif len(xv) == 1 and isinstance(xv[0], (tuple, list)) and len(xv[0]) == 2:
            q, p = xv[0]
        elif len(xv) == 1 and hasattr(xv[0], "shape") and xv[0].shape[-1] == 6:
            q, p = xv[0][..., :3], xv[0][..., 3:]
        else:
            raise NotImplementedError("HamiltonianField: expected (t, q, p), (t, (q, p)) or (t, qp[..., 6]); field args are not supported")
        return p, self.potential.acceleration(q, t)


def _as_potential(obj) -> AbstractPotential:
    if isinstance(obj, HamiltonianField):
        return obj.potential
    if isinstance(obj, AbstractPotential):
        return obj
    raise TypeError(f"expected a potential or HamiltonianField, got {type(obj).__name__}")


# ------------------------------------------------------------------------------------------------
# low-level launch wrapper


def _split_w0(w0):
    """-> q, p, t (t may be None).  Accepts what legacy/funcs.py:45 accepts."""
    if isinstance(w0, (PhaseSpaceCoordinate, Orbit)):
        return w0.q, w0.p, w0.t
    if isinstance(w0, PhaseSpacePosition):
        return w0.q, w0.p, None
    if isinstance(w0, (tuple, list)) and len(w0) == 2:
        return w0[0], w0[1], None
    if hasattr(w0, "shape") and w0.shape[-1] == 6:
        return w0[..., 0:3], w0[..., 3:6], None
    raise TypeError("w0 must be a PhaseSpaceCoordinate/Position, a (q, p) tuple or a (*batch, 6) array")


def _period_order(q, p, t0v, t1, torch):
    """Processing order for the adaptive kernel: most expensive particles first, neighbours alike.

    Cost proxy = (integration length) / (dynamical time r/|v|); cheap to compute, good enough to keep
    the lanes of a warp in step (north_star: "period-sorted particle ordering")."""
    r = q.norm(dim=-1)
    v = p.norm(dim=-1).clamp_min(1e-300)
    span = (t1 - t0v).abs() if t0v is not None else 1.0
    cost = span * v / r.clamp_min(1e-300)
    return torch.argsort(cost, descending=True).to(torch.int32).contiguous()


# Large batches that arrive in host memory (pinned tensors or numpy arrays) are processed in PIPELINE_CHUNKS slices on two
# side streams, so that
# the host->device copy of slice k+1 and the device->host copy of slice k-1 overlap the kernel of slice k (particles
# are independent; the kernels of neighbouring slices also fill each other's tail waves).  Same numbers as one launch.
PIPELINE_MIN_PARTICLES = 1 << 19
PIPELINE_CHUNKS = 4  # at least; up to PIPELINE_MAX_CHUNKS slices of ~PIPELINE_CHUNK_PARTICLES (1.2e6 particles: 8 slices, +1.2 %)
PIPELINE_MAX_CHUNKS = 8
PIPELINE_CHUNK_PARTICLES = 150_000
PIPELINE_MAX_OUTPUT_BYTES = 8 << 30


def _pipeline_ok(torch, q0, p0, t0, ts, layout) -> bool:
    if layout != "NT3":
        return False
    both_np = isinstance(q0, np.ndarray) and isinstance(p0, np.ndarray)
    if both_np:
        # numpy callers: pageable memory.  A pageable host->device copy blocks the host but not the kernels already
        # queued, so slice k+1 is copied while slice k runs -- same pipeline, only the first slice's copy is exposed.
        if q0.dtype != np.float64 or p0.dtype != np.float64 or q0.shape != p0.shape or q0.shape[-1] != 3:
            return False
        if not (q0.flags.c_contiguous and p0.flags.c_contiguous):
            return False
    else:
        if not (isinstance(q0, torch.Tensor) and isinstance(p0, torch.Tensor)) or q0.is_cuda or p0.is_cuda:
            return False
        if q0.dtype != torch.float64 or p0.dtype != torch.float64 or q0.shape != p0.shape or q0.shape[-1] != 3:
            return False
        if not (q0.is_contiguous() and p0.is_contiguous()):
            return False
    if isinstance(t0, torch.Tensor) and t0.is_cuda:
        return False
    n = int(np.prod(q0.shape[:-1]))
    n_saves = int(ts.numel()) if isinstance(ts, torch.Tensor) else int(np.size(ts))
    if n * max(n_saves, 1) * 48 > PIPELINE_MAX_OUTPUT_BYTES:  # (pinned host buffers for the whole result)
        return False
    return n >= PIPELINE_MIN_PARTICLES


_PIPELINE_STREAMS: dict[int, list] = {}


def _pipeline_streams(torch, dev):
    """Two side streams per device, created once: the caching allocators keep their blocks per stream, so fresh
    streams on every call would turn each call's buffers into new cudaMalloc / cudaHostAlloc calls."""
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _PIPELINE_STREAMS:
        _PIPELINE_STREAMS[key] = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    return _PIPELINE_STREAMS[key]


def _integrate_pipelined(pot, q0, p0, t0, t1, ts, *, solver, controller, dt0, max_steps, sort, throw, general_kernel):
    import torch

    from_numpy = isinstance(q0, np.ndarray)
    if from_numpy:
        q0, p0 = torch.from_numpy(q0), torch.from_numpy(p0)  # (views, no copy)
    batch = tuple(q0.shape[:-1])
    qh, ph = q0.reshape(-1, 3), p0.reshape(-1, 3)
    N = qh.shape[0]
    t0_is_array = (isinstance(t0, (np.ndarray, torch.Tensor)) and np.ndim(t0) > 0) or isinstance(t0, (list, tuple))
    t0h = None
    if t0_is_array:
        t0h = torch.as_tensor(np.asarray(t0, dtype=np.float64) if not isinstance(t0, torch.Tensor) else t0).reshape(-1)
        if t0h.shape[0] != N:
            raise ValueError("per-particle t0 must match the batch size")
    ts_np = np.atleast_1d(np.asarray(ts.detach().cpu() if isinstance(ts, torch.Tensor) else ts, dtype=np.float64))
    T = int(ts_np.shape[0])
    q_out = torch.empty((N, T, 3), dtype=torch.float64, pin_memory=True)
    p_out = torch.empty((N, T, 3), dtype=torch.float64, pin_memory=True)
    status = torch.empty((N,), dtype=torch.int32, pin_memory=True)
    stat_out: dict[str, Any] = {}
    dev = torch.device("cuda", torch.cuda.current_device())
    cur = torch.cuda.current_stream(dev)
    streams = _pipeline_streams(torch, dev)
    chunks = max(PIPELINE_CHUNKS, min(PIPELINE_MAX_CHUNKS, N // PIPELINE_CHUNK_PARTICLES))
    bounds = [(N * k) // chunks for k in range(chunks + 1)]
    for k in range(chunks):
        a, b = bounds[k], bounds[k + 1]
        s = streams[k % 2]
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            dq = qh[a:b].to(dev, non_blocking=True)
            dp = ph[a:b].to(dev, non_blocking=True)
            t0k = t0 if t0h is None else t0h[a:b].to(dev)
            qd, pd, st, stats = _integrate(pot, dq, dp, t0k, t1, ts_np, solver=solver, controller=controller, dt0=dt0,
                                           max_steps=max_steps, sort=sort, layout="NT3", throw=False,
                                           general_kernel=general_kernel)
            q_out[a:b].copy_(qd, non_blocking=True)
            p_out[a:b].copy_(pd, non_blocking=True)
            status[a:b].copy_(st, non_blocking=True)
            for name, v in stats.items():
                if name not in stat_out:
                    stat_out[name] = torch.empty((N,), dtype=v.dtype, pin_memory=True)
                stat_out[name][a:b].copy_(v, non_blocking=True)
    for s in streams:
        s.synchronize()
    if throw:
        bad = torch.nonzero(status != _lib.OK)
        if bad.numel():
            i = int(bad[0, 0])
            code = int(status[i])
            why = {1: "max_steps reached", 2: "non-finite state"}.get(code, f"status {code}")
            raise RuntimeError(f"integration failed for particle {i} of {N}: {why} ({bad.shape[0]} failed in total)")
    q_out, p_out = q_out.reshape(batch + (T, 3)), p_out.reshape(batch + (T, 3))
    if from_numpy:  # numpy in -> numpy out, as the single-launch path does
        q_out, p_out = q_out.numpy(), p_out.numpy()
    return q_out, p_out, status.reshape(batch), {k: v.reshape(batch) for k, v in stat_out.items()}


DIAGNOSTICS = ("energy", "angular_momentum", "tidal_tensor")
FUSE_MAX_SAVES_ADAPTIVE = 16


def _integrate(pot, q0, p0, t0, t1, ts, *, solver, controller, dt0, max_steps, sort=True, layout="NT3",
               throw=True, general_kernel=False, diagnostics=(), joint=False, fuse="auto"):  # fmt: skip
    """One launch of the integrator kernels.  q0, p0: (*batch, 3); t0 scalar or (*batch,); ts: (T,).

    ``joint=True`` (adaptive solvers, scalar t0): the whole batch is ONE ODE with one shared step and an error norm over
    all 6N components -- what the reference's scalar-time call forms do (orbit/solver.py:774-803); the default controls
    the step per particle (the reference under ``vmap``).  status and step counts are then the same for every particle.

    ``diagnostics``: any of ``"energy"``, ``"angular_momentum"``, ``"tidal_tensor"`` -- evaluated inside the integrator
    kernel at every saved state (``gx_integrate_*_epilogue``) and returned in the stats dict under those names with
    shapes (*batch, T), (*batch, T, 3), (*batch, T, 3, 3) (layout "T3N": (T, N), (T, 3, N), (T, 9, N)).
    ``fuse``: True = always in the kernel; False = a second pass over the saved states (same bits for E and L);
    "auto" (default) = in the kernel for the fixed-step schemes and for adaptive solves with at most
    ``FUSE_MAX_SAVES_ADAPTIVE`` saves.  Measured on a quarter of C2 (303 104 particles x 1000 saves, Dopri8): integration
    0.110 s, second pass 0.006 s (HBM speed, every lane busy), fused 0.235 s -- inside the adaptive kernel a save is
    evaluated by the ~5 lanes of a warp whose step contains one, so the potential evaluation is paid at a sixth of the
    lanes; the fixed-step kernels save rarely relative to their steps and fuse for free."""
    torch = _lib.require_cuda()
    diagnostics = tuple(diagnostics or ())
    for d in diagnostics:
        if d not in DIAGNOSTICS:
            raise ValueError(f"unknown diagnostic {d!r}; choose from {DIAGNOSTICS}")
    if joint and diagnostics:
        raise NotImplementedError("diagnostics= is not available with joint=True")
    if diagnostics and layout == "NT3" and (fuse is False or (fuse == "auto" and isinstance(solver, (Dopri8, Dopri5))
                                                              and np.size(ts) > FUSE_MAX_SAVES_ADAPTIVE)):
        q, p, status, stats = _integrate(pot, q0, p0, t0, t1, ts, solver=solver, controller=controller, dt0=dt0,
                                         max_steps=max_steps, sort=sort, layout=layout, throw=throw,
                                         general_kernel=general_kernel)
        for d in diagnostics:  # the second pass: one HBM-bound kernel per quantity over the saved states
            stats[d] = (_energy(pot, q, p) if d == "energy" else _energy(None, q, p, want="L") if d == "angular_momentum"
                        else pot.tidal_tensor(q, 0.0))
        return q, p, status, stats
    if not diagnostics and not joint and _pipeline_ok(torch, q0, p0, t0, ts, layout):
        return _integrate_pipelined(pot, q0, p0, t0, t1, ts, solver=solver, controller=controller, dt0=dt0,
                                    max_steps=max_steps, sort=sort, throw=throw, general_kernel=general_kernel)
    dq, restore = _to_device(q0)
    dp, _ = _to_device(p0)
    batch = tuple(dq.shape[:-1])
    dq = dq.reshape(-1, 3).contiguous()
    dp = dp.expand(*batch, 3).reshape(-1, 3).contiguous() if dp.shape != dq.shape else dp.reshape(-1, 3).contiguous()
    N = dq.shape[0]
    dev = dq.device
    ts_np = np.atleast_1d(np.asarray(ts.detach().cpu() if isinstance(ts, torch.Tensor) else ts, dtype=np.float64))
    T = int(ts_np.shape[0])
    dts = torch.from_numpy(ts_np).to(dev)
    t1 = float(t1)
    t0_arr = None
    t0_is_array = (isinstance(t0, (np.ndarray, torch.Tensor)) and np.ndim(t0) > 0) or isinstance(t0, (list, tuple))
    if t0_is_array:
        t0_arr, _ = _to_device(t0)
        t0_arr = t0_arr.reshape(-1).contiguous()
        if t0_arr.shape[0] != N:
            raise ValueError("per-particle t0 must match the batch size")
        t0s = 0.0
    else:
        t0s = float(t0)
    if layout == "NT3":
        q = torch.empty((N, T, 3), dtype=torch.float64, device=dev)
        p = torch.empty((N, T, 3), dtype=torch.float64, device=dev)
        lay = _lib.LAYOUT_NT3
    elif layout == "T3N":
        q = torch.empty((T, 3, N), dtype=torch.float64, device=dev)
        p = torch.empty((T, 3, N), dtype=torch.float64, device=dev)
        lay = _lib.LAYOUT_T3N
    else:
        raise ValueError("layout must be 'NT3' or 'T3N'")
    status = torch.empty((N,), dtype=torch.int32, device=dev)
    P = pot.c_struct()
    ms = -1 if max_steps is None else int(max_steps)
    stats: dict[str, Any] = {}
    L = _lib.lib()
    epi = None
    diag: dict[str, Any] = {}
    if diagnostics:
        nt3 = layout == "NT3"
        shapes = {"energy": (N, T) if nt3 else (T, N), "angular_momentum": (N, T, 3) if nt3 else (T, 3, N),
                  "tidal_tensor": (N, T, 3, 3) if nt3 else (T, 9, N)}
        for d in diagnostics:
            diag[d] = torch.empty(shapes[d], dtype=torch.float64, device=dev)
        ptr = lambda d: diag[d].data_ptr() if d in diag else None  # noqa: E731
        epi = _lib.GxOrbitEpilogue(ptr("energy"), ptr("angular_momentum"), ptr("tidal_tensor"))
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        if isinstance(solver, (SemiImplicitEuler, LeapfrogMidpoint)):
            if not isinstance(controller, ConstantStepSize):
                raise NotImplementedError(f"{type(solver).__name__} requires ConstantStepSize()")
            joint = False  # (a constant step is shared by construction: both call forms are the same computation)
            if dt0 is None:
                raise ValueError("ConstantStepSize requires dt0")
            if t0_arr is not None:
                raise NotImplementedError("per-particle t0 is only supported by the adaptive solver")
            scheme = _lib.SCHEME_SIE if isinstance(solver, SemiImplicitEuler) else _lib.SCHEME_LEAPFROG_MIDPOINT
            if general_kernel:  # testing aid: the per-step time arithmetic kernel instead of the run-length one
                scheme |= _lib.SCHEME_GENERAL_KERNEL
            if solver.strict:
                scheme |= _lib.SCHEME_STRICT
            if epi is not None:
                rc = L.gx_integrate_fixed_epilogue(C.byref(P), dq.data_ptr(), dp.data_ptr(), N, t0s, t1, float(dt0),
                                                   dts.data_ptr(), T, scheme, ms, lay, q.data_ptr(), p.data_ptr(),
                                                   status.data_ptr(), C.byref(epi), stream)  # fmt: skip
            else:
                rc = L.gx_integrate_fixed(C.byref(P), dq.data_ptr(), dp.data_ptr(), N, t0s, t1, float(dt0),
                                          dts.data_ptr(), T, scheme, ms, lay, q.data_ptr(), p.data_ptr(),
                                          status.data_ptr(), stream)  # fmt: skip
            _lib.check(rc, "gx_integrate_fixed")
        elif isinstance(solver, (Dopri8, Dopri5)):
            if not isinstance(controller, PIDController):
                raise NotImplementedError("Dopri8 / Dopri5 require a PIDController")
            code = _lib.SOLVER_DOPRI8 if isinstance(solver, Dopri8) else _lib.SOLVER_DOPRI5
            if solver.strict:
                code |= _lib.SOLVER_STRICT
            pid = controller.c_struct(dt0)
            if joint:
                if t0_arr is not None or solver.strict:
                    raise NotImplementedError("joint=True needs a scalar t0 (batched start times are per-particle "
                                              "solves in the reference too) and is not available in strict mode")
                one = torch.empty((3,), dtype=torch.int32, device=dev)
                ws = torch.empty((int(L.gx_joint_workspace_bytes(N)) // 8,), dtype=torch.float64, device=dev)
                rc = L.gx_integrate_adaptive_joint(code, C.byref(P), C.byref(pid), dq.data_ptr(), dp.data_ptr(), N, t0s, t1,
                                                   dts.data_ptr(), T, ms, lay, q.data_ptr(), p.data_ptr(),
                                                   one[0:].data_ptr(), one[1:].data_ptr(), one[2:].data_ptr(),
                                                   ws.data_ptr(), stream)  # fmt: skip
                _lib.check(rc, "gx_integrate_adaptive_joint")
                status = one[0:1].expand(N).contiguous()
                nacc, ntot = one[1:2].expand(N).contiguous(), one[2:3].expand(N).contiguous()
                order = None
            else:
                nacc = torch.empty((N,), dtype=torch.int32, device=dev)
                ntot = torch.empty((N,), dtype=torch.int32, device=dev)
                ws = torch.empty((int(L.gx_workspace_bytes()) // 8,), dtype=torch.int64, device=dev)
                order = _period_order(dq, dp, t0_arr, t1, torch) if (sort and N > 64) else None
            if joint:
                pass
            elif epi is not None:
                rc = L.gx_integrate_adaptive_epilogue(
                    code, C.byref(P), C.byref(pid), dq.data_ptr(), dp.data_ptr(), N,
                    None if t0_arr is None else t0_arr.data_ptr(), t0s, t1, dts.data_ptr(), T, ms,
                    None if order is None else order.data_ptr(), lay, q.data_ptr(), p.data_ptr(), status.data_ptr(),
                    nacc.data_ptr(), ntot.data_ptr(), ws.data_ptr(), C.byref(epi), stream)  # fmt: skip
                _lib.check(rc, "gx_integrate_adaptive_epilogue")
            elif N == 1 and T >= 64 and t0_arr is None and t0s != t1 and layout == "NT3" and not solver.strict:
                # single orbit, many saves (mock-stream progenitor): record the steps, evaluate the dense output
                # for all save times in parallel (gx_integrate_dopri8_record + gx_dense_eval)
                cap = int(min(ms, 1 << 16)) if ms > 0 else (1 << 16)
                rec = torch.empty((cap, _lib.DENSE_RECORD_DOUBLES), dtype=torch.float64, device=dev)
                n_rec = torch.zeros((1,), dtype=torch.int32, device=dev)
                rc = L.gx_integrate_adaptive_record(code, C.byref(P), C.byref(pid), dq.data_ptr(), dp.data_ptr(), t0s,
                                                    t1, ms, rec.data_ptr(), cap, n_rec.data_ptr(), status.data_ptr(),
                                                    nacc.data_ptr(), ntot.data_ptr(), ws.data_ptr(), stream)  # fmt: skip
                _lib.check(rc, "gx_integrate_adaptive_record")
                rc = L.gx_dense_eval_solver(code, rec.data_ptr(), n_rec.data_ptr(), t0s, t1, dts.data_ptr(), T,
                                            q.data_ptr(), p.data_ptr(), stream)  # fmt: skip
                _lib.check(rc, "gx_dense_eval_solver")
            else:
              rc = L.gx_integrate_adaptive(code, C.byref(P), C.byref(pid), dq.data_ptr(), dp.data_ptr(), N,
                                       None if t0_arr is None else t0_arr.data_ptr(), t0s, t1, dts.data_ptr(), T, ms,
                                       None if order is None else order.data_ptr(), lay, q.data_ptr(), p.data_ptr(),
                                       status.data_ptr(), nacc.data_ptr(), ntot.data_ptr(), ws.data_ptr(),
                                       stream)  # fmt: skip
              _lib.check(rc, "gx_integrate_adaptive")
            stats["num_accepted_steps"] = nacc.reshape(batch)
            stats["num_steps"] = ntot.reshape(batch)
        else:
            raise NotImplementedError(
                f"solver {type(solver).__name__} is not supported (SemiImplicitEuler, LeapfrogMidpoint, Dopri8, Dopri5)"
            )
    if layout == "NT3":
        q = q.reshape(batch + (T, 3))
        p = p.reshape(batch + (T, 3))
        for d, tail in (("energy", (T,)), ("angular_momentum", (T, 3)), ("tidal_tensor", (T, 3, 3))):
            if d in diag:
                diag[d] = diag[d].reshape(batch + tail)
    stats.update(diag)
    host_caller = not (hasattr(q0, "is_cuda") and q0.is_cuda)
    if host_caller:
        # Host callers get host results and host bookkeeping (int32 torch tensors on the CPU: np.asarray(...) /
        # .numpy() work, no device round trip later).  Everything goes back through pinned memory in one batch of
        # asynchronous copies with a single synchronisation; the status check below then reads host memory.
        q, p, status, *rest = _to_host_many([q, p, status, *stats.values()])
        stats = dict(zip(stats.keys(), rest))
        if isinstance(q0, np.ndarray) or not isinstance(q0, torch.Tensor):
            q, p = q.numpy(), p.numpy()
            for d in diag:
                stats[d] = stats[d].numpy()
    if throw:
        bad = torch.nonzero(status != _lib.OK)
        if bad.numel():
            i = int(bad[0, 0])
            code = int(status[i])
            why = {1: "max_steps reached", 2: "non-finite state"}.get(code, f"status {code}")
            raise RuntimeError(f"integration failed for particle {i} of {N}: {why} ({bad.shape[0]} failed in total)")
    return q, p, status.reshape(batch), stats


def _energy(pot, q, p, want="E", t=None):
    """E = |p|^2/2 + Phi(q) (``pot`` None: kinetic energy only) or L = q x p, on the device.

    ``t``: the time(s) of the states, needed for potentials with LinearParameters -- a scalar, or the (T,) save times of
    an orbit batch (*batch, T, 3)."""
    torch = _lib.require_cuda()
    dq, restore = _to_device(q)
    dp, _ = _to_device(p)
    batch = tuple(dq.shape[:-1])
    dq = dq.reshape(-1, 3).contiguous()
    dp = dp.reshape(-1, 3).contiguous()
    n = dq.shape[0]
    E = torch.empty((n,), dtype=torch.float64, device=dq.device) if want == "E" else None
    L = torch.empty((n, 3), dtype=torch.float64, device=dq.device) if want == "L" else None
    P = pot.c_struct() if pot is not None else NullPotentialStruct()
    with torch.cuda.device(dq.device):
        if t is not None and pot is not None and want == "E":
            tn = np.asarray(t.detach().cpu() if hasattr(t, "detach") else t, dtype=np.float64)
            if tn.ndim == 0:
                tv, period, tsc = None, 1, float(tn)
            else:
                if batch[-1:] != tn.shape:
                    raise ValueError(f"times of shape {tn.shape} do not match states of shape {batch}")
                tv, period, tsc = torch.from_numpy(np.ascontiguousarray(tn)).to(dq.device), int(tn.shape[0]), 0.0
            rc = _lib.lib().gx_energy_angmom_t(C.byref(P), dq.data_ptr(), dp.data_ptr(), n,
                                               None if tv is None else tv.data_ptr(), period, tsc, E.data_ptr(), None,
                                               torch.cuda.current_stream().cuda_stream)  # fmt: skip
        else:
            rc = _lib.lib().gx_energy_angmom(C.byref(P), dq.data_ptr(), dp.data_ptr(), n,
                                             None if E is None else E.data_ptr(), None if L is None else L.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream)  # fmt: skip
    _lib.check(rc, "gx_energy_angmom")
    return restore(E.reshape(batch)) if want == "E" else restore(L.reshape(batch + (3,)))


def NullPotentialStruct() -> _lib.GxPotential:
    """A potential with no components (Phi = 0)."""
    P = _lib.GxPotential()
    P.n = 0
    P.G = 1.0
    return P


# ------------------------------------------------------------------------------------------------
# OrbitSolver / Integrator


@dataclasses.dataclass(frozen=True)
class OrbitSolver:
    """orbit/solver.py:121-141.  Defaults: Dopri8, PIDController(1e-8, 1e-8), max_steps=2**16."""

    solver: Any = dataclasses.field(default_factory=Dopri8)
    stepsize_controller: Any = dataclasses.field(default_factory=lambda: PIDController(rtol=1e-8, atol=1e-8))
    max_steps: int | None = 2**16
    event: Any = None

    def solve(self, field, w0, t0, t1=None, /, *, saveat=None, dt0=None, max_steps="default", args=None,
              dense=False, unbatch_time=False, throw=True, sort=True, diagnostics=(), joint=False):  # fmt: skip
        """``OrbitSolver.solve(field, (q0, p0), t0, t1, saveat=ts, dt0=..., max_steps=...)``.

        Like orbit/solver.py:774-803; when ``w0`` carries a time, the 3-argument form
        ``solve(field, w0, t1)`` is accepted (orbit/solver.py:431-442).

        Batch semantics.  The reference solves a batch ``(q0[N,3], p0[N,3])`` with scalar times as ONE ODE with a shared
        adaptive step (and per particle under ``lstrat.VMap`` / batched times).  Here the default is per particle for
        every call form (north_star: per-particle step control); ``joint=True`` selects the reference's shared-step
        form for scalar times.  The two agree to the tolerance, not to rounding.
        """
        if args is not None or dense or self.event is not None:
            raise NotImplementedError("args / dense=True / events are not supported by the CUDA path")
        pot = _as_potential(field)
        q0, p0, tw = _split_w0(w0)
        if t1 is None:
            if tw is None:
                raise ValueError("solve(field, w0, t1) needs w0 to carry a time")
            t0, t1 = tw, t0
        t1f = float(np.asarray(t1))
        ts = np.atleast_1d(np.asarray([t1f] if saveat is None else saveat, dtype=np.float64))
        ms = self.max_steps if max_steps == "default" else max_steps
        q, p, status, stats = _integrate(pot, q0, p0, t0, t1f, ts, solver=self.solver,
                                         controller=self.stepsize_controller, dt0=dt0, max_steps=ms, throw=throw,
                                         sort=sort, diagnostics=diagnostics, joint=joint)  # fmt: skip
        if unbatch_time and ts.shape[0] == 1:
            q, p = q[..., 0, :], p[..., 0, :]
        return Solution(t0=t0, t1=t1f, ts=ts, ys=(q, p), stats=stats, result=status)


def _default_integrator_solver() -> OrbitSolver:
    # legacy/integrator.py:37-39,161-168: Dopri8, rtol = atol = 1e-7
    return OrbitSolver(solver=Dopri8(), stepsize_controller=PIDController(rtol=1e-7, atol=1e-7))


@dataclasses.dataclass(frozen=True)
class Integrator:
    """legacy/integrator.py:42-245.  ``diffeq_kw`` defaults to ``max_steps=None`` (:170-174)."""

    dynamics_solver: OrbitSolver = dataclasses.field(default_factory=_default_integrator_solver)
    diffeq_kw: dict = dataclasses.field(default_factory=lambda: {"max_steps": None})

    def __call__(self, field, w0, t0, t1, /, *, saveat=None, dense=False, throw=True, diagnostics=(), joint=False):
        if dense:
            raise NotImplementedError("dense=True (interpolated orbits) is not supported by the CUDA path")
        kw = dict(self.diffeq_kw)
        q0, p0, _ = _split_w0(w0)
        sol = self.dynamics_solver.solve(field, (q0, p0), t0, t1, saveat=saveat, dt0=kw.get("dt0"),
                                         max_steps=kw.get("max_steps", "default"), throw=throw,
                                         diagnostics=diagnostics, joint=joint)  # fmt: skip
        q, p = sol.ys
        if saveat is None:
            return PhaseSpaceCoordinate(q[..., 0, :], p[..., 0, :], sol.t1)
        w = PhaseSpaceCoordinate(q, p, sol.ts)
        if diagnostics:
            w.diagnostics = {d: sol.stats[d] for d in diagnostics}
        return w


default_integrator = Integrator()


def evaluate_orbit(pot, w0, t, /, *, integrator: Integrator | None = None, dense: bool = False, throw=True,
                   diagnostics=(), joint=False) -> Orbit:  # fmt: skip
    """``gd.evaluate_orbit`` (legacy/funcs.py:42-213): integrate w0 to t[0], then t[0] -> t[-1] saving at t.

    ``diagnostics`` (an extension): any of "energy", "angular_momentum", "tidal_tensor" -- evaluated by the integrator
    kernel at every saved state; ``Orbit.total_energy()`` / ``.angular_momentum()`` / ``.tidal_tensor()`` then return
    them without a second pass over the orbit.

    ``joint=True``: the reference's batch semantics for this call form -- the batch is one ODE with a shared adaptive
    step (legacy/integrator.py:288-298); the default controls the step per particle (see ``OrbitSolver.solve``)."""
    if dense:
        raise NotImplementedError("dense=True is not supported by the CUDA path")
    pot = _as_potential(pot)
    integrator = default_integrator if integrator is None else integrator
    t_host = np.atleast_1d(np.asarray(t.detach().cpu() if hasattr(t, "detach") else t, dtype=np.float64))
    q0, p0, tw0 = _split_w0(w0)
    field = HamiltonianField(pot)
    if tw0 is not None:
        # integration A: w0.t -> t[0] (legacy/funcs.py:194-206); zero-length when equal
        tw0f = np.asarray(tw0.detach().cpu() if hasattr(tw0, "detach") else tw0, dtype=np.float64)
        if tw0f.ndim == 0 and float(tw0f) == float(t_host[0]):
            pass
        else:
            w = integrator(field, (q0, p0), tw0 if tw0f.ndim else float(tw0f), float(t_host[0]), throw=throw,
                           joint=joint and not tw0f.ndim)
            q0, p0 = w.q, w.p
    # integration B: t[0] -> t[-1], saveat = t (legacy/funcs.py:210)
    w = integrator(field, (q0, p0), float(t_host[0]), float(t_host[-1]), saveat=t_host, throw=throw,
                   diagnostics=diagnostics, joint=joint)
    return Orbit(q=w.q, p=w.p, t=t_host, potential=pot, diagnostics=getattr(w, "diagnostics", None))


def compute_orbit(pot_or_field, w0, ts, /, *, solver: OrbitSolver | None = None, dense: bool = False,
                  throw=True, diagnostics=(), joint=False) -> Orbit:  # fmt: skip
    """``gd.compute_orbit`` (orbit/compute.py:28-98): same two-phase solve with an ``OrbitSolver``."""
    if dense:
        raise NotImplementedError("dense=True is not supported by the CUDA path")
    pot = _as_potential(pot_or_field)
    solver = OrbitSolver() if solver is None else solver
    t_host = np.atleast_1d(np.asarray(ts.detach().cpu() if hasattr(ts, "detach") else ts, dtype=np.float64))
    q0, p0, tw0 = _split_w0(w0)
    if tw0 is not None and float(np.asarray(tw0)) != float(t_host[0]):
        s0 = solver.solve(pot, (q0, p0), float(np.asarray(tw0)), float(t_host[0]), throw=throw, joint=joint)
        q0, p0 = s0.ys[0][..., 0, :], s0.ys[1][..., 0, :]
    sol = solver.solve(pot, (q0, p0), float(t_host[0]), float(t_host[-1]), saveat=t_host, throw=throw,
                       diagnostics=diagnostics, joint=joint)
    return Orbit(q=sol.ys[0], p=sol.ys[1], t=t_host, potential=pot, status=sol.result, n_steps=sol.stats,
                 diagnostics={d: sol.stats[d] for d in diagnostics} if diagnostics else None)


# ------------------------------------------------------------------------------------------------
# mock streams


@dataclasses.dataclass
class MockStreamArm:
    """dynamics/_src/mockstream/arm.py:20-98."""

    q: Any
    p: Any
    t: Any
    release_time: Any

    def w(self):
        return _cat(self.q, self.p)


class MockStream(dict):
    """dynamics/_src/mockstream/core.py:20-84: {'lead', 'trail'}; concatenated views sorted by release time."""

    def _joined(self, name):
        arms = list(self.values())
        import torch

        parts = [getattr(a, name) for a in arms]
        rel = [a.release_time for a in arms]
        if isinstance(parts[0], np.ndarray):
            order = np.argsort(np.concatenate(rel), kind="stable")
            return np.concatenate(parts)[order]
        order = torch.argsort(torch.cat(rel), stable=True)
        return torch.cat(parts)[order]

    @property
    def q(self):
        return self._joined("q")

    @property
    def p(self):
        return self._joined("p")

    @property
    def release_time(self):
        return self._joined("release_time")


class AbstractStreamDF:
    """legacy/mockstream/df/base.py:28-131.  ``rng`` supplies the random draws (see ``_draws``)."""

    df_kind: int = -1

    def _draws(self, rng, M: int) -> np.ndarray:
        raise NotImplementedError

    def sample(self, rng, pot, prog_orbit: Orbit, prog_mass):
        """-> dict(lead=MockStreamArm, trail=MockStreamArm) of release conditions along ``prog_orbit``."""
        torch = _lib.require_cuda()
        pot = _as_potential(pot)
        dq, restore = _to_device(prog_orbit.q)
        dp, _ = _to_device(prog_orbit.p)
        dq = dq.reshape(-1, 3).contiguous()
        dp = dp.reshape(-1, 3).contiguous()
        M = dq.shape[0]
        ts = prog_orbit.t
        if callable(prog_mass):  # ProgenitorMassCallable (df/progenitor.py:30-50)
            mass = np.asarray(prog_mass(np.asarray(ts, dtype=np.float64)), dtype=np.float64)
        else:
            mass = np.full((M,), float(getattr(prog_mass, "value", prog_mass)))
        dm = torch.from_numpy(np.ascontiguousarray(mass)).to(dq.device)
        is_key = isinstance(rng, np.ndarray) and rng.dtype == np.uint32 and rng.shape == (2,)  # raw jax key data
        draws = rng if (isinstance(rng, (np.ndarray, torch.Tensor)) and not is_key) else self._draws(rng, M)
        want = (4, M) if self.df_kind == _lib.DF_FARDAL15 else (M, 6)
        if tuple(draws.shape) != want:
            raise ValueError(f"random draws must have shape {want} for {type(self).__name__}, got {tuple(draws.shape)}")
        dd, _ = _to_device(draws)
        dd = dd.contiguous()
        outs = [torch.empty((M, 3), dtype=torch.float64, device=dq.device) for _ in range(4)]
        P = pot.c_struct()
        # every stripping time sees the potential of its own time (fardal15.py:49-94: tidal_radius(..., t=t)); for a
        # static potential the times are ignored
        tt, _ = _to_device(np.broadcast_to(np.asarray(ts, dtype=np.float64), (M,)).copy() if not hasattr(ts, "is_cuda") else ts)
        tt = tt.reshape(-1).contiguous()
        with torch.cuda.device(dq.device):
            rc = _lib.lib().gx_stream_release_t(C.byref(P), self.df_kind, dq.data_ptr(), dp.data_ptr(), dm.data_ptr(),
                                                tt.data_ptr(), dd.data_ptr(), M, *[o.data_ptr() for o in outs],
                                                torch.cuda.current_stream().cuda_stream)  # fmt: skip
        _lib.check(rc, "gx_stream_release_t")
        ql, pl, qt, pt = (restore(o) for o in outs)
        return {"lead": MockStreamArm(ql, pl, ts, ts), "trail": MockStreamArm(qt, pt, ts, ts)}


def _device_jax_normal(key_data, n: int):
    """``jax.random.normal(key, (n,))`` on the device (``gx_jax_normal``; threefry2x32, partitionable) -> CUDA tensor."""
    torch = _lib.require_cuda()
    out = torch.empty((n,), dtype=torch.float64, device="cuda")
    rc = _lib.lib().gx_jax_normal(int(key_data[0]), int(key_data[1]), n, out.data_ptr(),
                                  torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "gx_jax_normal")
    return out


class FardalStreamDF(AbstractStreamDF):
    """df/fardal15.py: k_r = 2 + 0.5 n1, k_vphi = k_r (0.3 + 0.5 n2), k_z = 0.5 n3, k_vz = 0.5 n4."""

    df_kind = _lib.DF_FARDAL15

    def _draws(self, rng, M):
        """An integer seed (or raw uint32[2] key data) reproduces ``jr.key(seed)`` -> ``jr.split(key, 4)`` ->
        ``jr.normal(key_i, (M, 1))`` of df/fardal15.py:61,81-84 (see galax_b200/jaxrandom.py); a
        ``numpy.random.Generator`` draws from numpy's stream instead."""
        if isinstance(rng, np.random.Generator):
            return rng.standard_normal((4, M))
        from . import jaxrandom

        torch = _lib.require_cuda()
        k = jaxrandom.key(rng) if np.ndim(rng) == 0 else np.asarray(rng, dtype=np.uint32)
        ks = jaxrandom.split(k, 4)  # four threefry evaluations on the host; the 4 M draws are made on the device
        return torch.stack([_device_jax_normal(ks[i], M) for i in range(4)])


class ChenStreamDF(AbstractStreamDF):
    """df/chen24.py: 6-D Gaussian in (Dr/r_t, phi, theta, Dv/v_esc, alpha, beta), 'no progenitor' version."""

    df_kind = _lib.DF_CHEN24
    mean = np.array([1.6, -30, 0, 1, 20, 0], dtype=np.float64)
    cov = np.array(
        [
            [0.1225, 0, 0, 0, -4.9, 0],
            [0, 529, 0, 0, 0, 0],
            [0, 0, 144, 0, 0, 0],
            [0, 0, 0, 0, 0, 0],
            [-4.9, 0, 0, 0, 400, 0],
            [0, 0, 0, 0, 0, 484],
        ],
        dtype=np.float64,
    )

    def _draws(self, rng, M):
        """An integer seed (or raw key data) follows ``jr.multivariate_normal(jr.key(seed), mean, cov, (M,),
        method="svd")`` of df/chen24.py:89-91 (threefry normals + LAPACK SVD, see jaxrandom.multivariate_normal_svd);
        a ``numpy.random.Generator`` draws from numpy's stream."""
        if isinstance(rng, np.random.Generator):
            return rng.multivariate_normal(self.mean, self.cov, size=M, method="svd")
        from . import jaxrandom

        torch = _lib.require_cuda()
        k = jaxrandom.key(rng) if np.ndim(rng) == 0 else np.asarray(rng, dtype=np.uint32)
        u, s, _ = np.linalg.svd(self.cov)
        factor = torch.from_numpy(u * np.sqrt(s)[None, :]).to("cuda")
        z = _device_jax_normal(k, 6 * M).reshape(M, 6)  # normal(key, (M, 6)), row-major counters
        return torch.from_numpy(self.mean).to("cuda") + z @ factor.T


@dataclasses.dataclass(frozen=True)
class MockStreamGenerator:
    """legacy/mockstream/mockstream_generator.py:33-275."""

    df: AbstractStreamDF
    potential: AbstractPotential
    progenitor_integrator: Integrator = dataclasses.field(default_factory=Integrator)
    stream_integrator: Integrator = dataclasses.field(default_factory=Integrator)

    def run(self, rng, ts, prog_w0, prog_mass, *, vmapped: bool | None = None, throw=True):
        """-> (MockStream, final progenitor PhaseSpaceCoordinate).

        ``rng``: an integer seed (Fardal: reproduces the reference's ``jr.key(seed)`` draws through the threefry
        restatement in ``galax_b200/jaxrandom.py``; Chen: the same threefry normals through numpy's LAPACK SVD), a
        ``numpy.random.Generator``, or the draws themselves (Fardal: (4, M) standard normals; Chen:
        (M, 6) samples).  ``vmapped`` is accepted for signature parity and ignored: every particle is an independent
        lane of the work queue.
        """
        ts = np.asarray(ts.detach().cpu() if hasattr(ts, "detach") else ts, dtype=np.float64)
        if ts.ndim != 1 or ts.shape[0] < 2:
            raise ValueError("ts must be a 1-D array of at least two stripping times")
        q0, p0, tw0 = _split_w0(prog_w0)
        if np.ndim(q0) != 1:
            raise ValueError("prog_w0 must be scalar")  # mockstream_generator.py:230
        if ts[1] < ts[0]:  # cond_reverse (mockstream_generator.py:236)
            ts = ts[::-1].copy()
        # the pipeline (progenitor orbit -> release -> stream integration) stays on the device; results go back to
        # the caller's array kind once, at the end
        _lib.require_cuda()
        q0d, to_caller = _to_device(q0)
        p0d, _ = _to_device(p0)
        w0 = PhaseSpaceCoordinate(q0d, p0d, ts[0] if tw0 is None else tw0)
        # progenitor orbit saved at the stripping times (:239)
        prog_o = evaluate_orbit(self.potential, w0, ts, integrator=self.progenitor_integrator, throw=throw)
        mock0 = self.df.sample(rng, self.potential, prog_o, prog_mass)
        # stream particles: release time -> t_f = ts[-1] + 1e-3, keep the final state (:88,139)
        t_f = float(ts[-1]) + 1e-3
        field = HamiltonianField(self.potential)
        # both arms in ONE launch: 2M independent particles with per-particle start times on the work queue
        M = ts.shape[0]
        lead, trail = mock0["lead"], mock0["trail"]
        q_all, p_all = _cat0(lead.q, trail.q), _cat0(lead.p, trail.p)
        w = self.stream_integrator(field, (q_all, p_all), np.concatenate([ts, ts]), t_f, throw=throw)
        tt = np.ones_like(ts) * ts[-1]
        wq, wp = to_caller(w.q), to_caller(w.p)
        arms = {"lead": MockStreamArm(wq[:M], wp[:M], tt, lead.release_time),
                "trail": MockStreamArm(wq[M:], wp[M:], tt, trail.release_time)}  # fmt: skip
        last = prog_o[-1]
        last = PhaseSpaceCoordinate(to_caller(last.q), to_caller(last.p), last.t)
        return MockStream(arms), last


__all__ = [
    "SemiImplicitEuler", "LeapfrogMidpoint", "Dopri8", "Dopri5", "ConstantStepSize", "PIDController",
    "PhaseSpacePosition", "PhaseSpaceCoordinate", "Orbit", "Solution", "HamiltonianField",
    "OrbitSolver", "Integrator", "default_integrator", "evaluate_orbit", "compute_orbit",
    "MockStreamArm", "MockStream", "AbstractStreamDF", "FardalStreamDF", "ChenStreamDF", "MockStreamGenerator",
]  # fmt: skip
