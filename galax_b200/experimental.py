"""Mirror of ``galax.dynamics.experimental``: the raw-array API the reference's own
benchmark suite targets (``/root/reference/tests/benchmark/test_experimental.py``).

* ``integrate_orbit([loop_strategy,] pot, (q0, p0), t0, t1, *, saveat=, solver=, solver_kwargs=)``
  -- ``dynamics/_src/experimental/integrate.py:322-575``; default solver Dopri8 with
  ``PIDController(rtol=1e-7, atol=1e-7, dtmin=0.3, force_dtmin=True)``, ``max_steps=10_000`` (``:28-35``).
* ``Fardal2015DF().sample(draws, pot, t, x, v, Msat)`` -- ``experimental/df.py:87-175`` (parameterised means / sigmas).
* ``StreamSimulator().init(...)`` / ``.run(...)`` -- ``experimental/stream.py:75-446``; default solver **Dopri5**,
  same controller (``:32-41``).

Loop strategies (``Scan`` / ``VMap`` / ``NoLoop``, ``dynamics/loop_strategies.py``) are accepted and ignored: every
orbit is an independent lane of the work queue, which is what all three compute.  Random draws: a seed /
``numpy.random.Generator`` or the standard-normal draws themselves; jax's key stream is not reproduced.
"""

from __future__ import annotations

import dataclasses
from typing import Any, Mapping

import numpy as np

from . import _lib
from .dynamics import (Dopri5, Dopri8, HamiltonianField, PIDController, _as_potential, _cat0, _integrate,
                       _to_device)  # fmt: skip
from .potential import AbstractPotential


class AbstractLoopStrategy:
    """dynamics/loop_strategies.py:8-47 (flags only)."""


class Determine(AbstractLoopStrategy):
    pass


class NoLoop(AbstractLoopStrategy):
    pass


class Scan(AbstractLoopStrategy):
    pass


class VMap(AbstractLoopStrategy):
    pass


class Vectorize(AbstractLoopStrategy):
    pass


@dataclasses.dataclass(frozen=True)
class DiffEqSolver:
    """Stand-in for ``diffraxtra.DiffEqSolver(solver, stepsize_controller, max_steps)``."""

    solver: Any
    stepsize_controller: Any
    max_steps: int | None = 10_000


def _default_controller() -> PIDController:
    return PIDController(rtol=1e-7, atol=1e-7, dtmin=0.3, dtmax=None, force_dtmin=True)


default_solver = DiffEqSolver(solver=Dopri8(), stepsize_controller=_default_controller(), max_steps=10_000)
default_stream_solver = DiffEqSolver(solver=Dopri5(), stepsize_controller=_default_controller(), max_steps=10_000)


@dataclasses.dataclass
class Solution:
    t0: Any
    t1: Any
    ts: Any
    ys: tuple
    stats: dict
    result: Any


def parse_t0_t1_saveat(t0, t1, saveat):
    """experimental/integrate.py:38-145: the three argument patterns."""
    if saveat is None:
        if t0 is None or t1 is None:
            raise ValueError("t0 and t1 must be specified if saveat is None")
        return float(t0), float(t1), None
    sv = np.asarray(saveat, dtype=np.float64)
    if sv.ndim == 0:
        if t0 is None:
            raise ValueError("t0 must be specified if saveat is a scalar")
        return float(t0), float(sv), np.array([float(sv)])
    if sv.shape[0] < 2:
        raise ValueError("saveat must be a scalar or have at least two elements")
    return (float(sv[0]) if t0 is None else float(t0)), (float(sv[-1]) if t1 is None else float(t1)), sv


def integrate_orbit(*args, t0=None, t1=None, saveat=None, solver: DiffEqSolver = default_solver,
                    solver_kwargs: Mapping[str, Any] | None = None, dense: bool = False,
                    dense_vectorize: bool = False, throw: bool = True) -> Solution:  # fmt: skip
    """``integrate_orbit([lstrat,] pot, (q0, p0), t0, t1, saveat=...)``.

    ``q0, p0``: ``(3,)`` or ``(B, 3)``; ``t0`` may be per-orbit ``(B,)`` (experimental/integrate.py:446-509).
    Returns a ``Solution`` with ``ys = (q, p)`` shaped ``(*batch, T, 3)`` (``(*batch, 3)`` when ``saveat`` is None or a
    scalar, like the reference's ``t1=True`` save).
    """
    if dense or dense_vectorize:
        raise NotImplementedError("dense interpolated solutions are not supported by the CUDA path")
    args = list(args)
    if args and isinstance(args[0], type) and issubclass(args[0], AbstractLoopStrategy):
        args.pop(0)
    if len(args) < 2:
        raise TypeError("integrate_orbit(pot, (q0, p0), t0, t1, ...)")
    pot = _as_potential(args[0])
    q0, p0 = args[1]
    if len(args) > 2:
        t0 = args[2]
    if len(args) > 3:
        t1 = args[3]
    kw = dict(solver_kwargs or {})
    ctrl = kw.get("stepsize_controller", solver.stepsize_controller)
    max_steps = kw.get("max_steps", solver.max_steps)
    t0_arr = None
    if t0 is not None and np.ndim(t0) > 0:
        t0_arr, t0 = t0, None
    t0s, t1s, sv = parse_t0_t1_saveat(0.0 if (t0 is None and t0_arr is not None) else t0, t1, saveat)
    ts = np.array([t1s]) if sv is None else sv
    q, p, status, stats = _integrate(pot, q0, p0, t0_arr if t0_arr is not None else t0s, t1s, ts, solver=solver.solver,
                                     controller=ctrl, dt0=kw.get("dt0"), max_steps=max_steps, throw=throw)  # fmt: skip
    if sv is None or np.ndim(saveat) == 0:
        q, p = q[..., 0, :], p[..., 0, :]
    return Solution(t0=t0_arr if t0_arr is not None else t0s, t1=t1s, ts=ts, ys=(q, p), stats=stats, result=status)


# ------------------------------------------------------------------------------------------------


def _is_key_data(key) -> bool:
    return isinstance(key, np.ndarray) and key.dtype == np.uint32 and key.shape == (2,)


def _fardal_normals(key, M: int):
    """(4, M) standard normals on the device: an int seed / uint32[2] key reproduces jax's stream (``jr.split(key, 4)``,
    each ``jr.normal(k_i, (M,))``, generated by ``gx_jax_normal``), a ``numpy.random.Generator`` uses numpy's, a float
    array / tensor is taken as the draws themselves."""
    import torch

    from . import jaxrandom
    from .dynamics import _device_jax_normal

    if isinstance(key, np.random.Generator):
        return torch.from_numpy(key.standard_normal((4, M))).to("cuda")
    if isinstance(key, (int, np.integer)) or _is_key_data(key):
        k = jaxrandom.key(key) if isinstance(key, (int, np.integer)) else np.asarray(key, dtype=np.uint32)
        ks = jaxrandom.split(k, 4)
        return torch.stack([_device_jax_normal(ks[i], M) for i in range(4)])
    if isinstance(key, torch.Tensor):
        return key.to(device="cuda", dtype=torch.float64).reshape(4, M)
    return torch.from_numpy(np.ascontiguousarray(np.asarray(key, dtype=np.float64).reshape(4, M))).to("cuda")


def _fardal_chain_normals(key_data, M: int):
    """(4, M) normals of the ``StreamSimulator.init`` scan (``key, subkey = jr.split(key)`` per release time, then four
    scalar normals on ``jr.split(subkey, 4)``): ``gx_jax_fardal_chain`` -- key chain and draws on the device."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    out = torch.empty((4, M), dtype=torch.float64, device="cuda")
    ws = torch.empty((max(int(L.gx_jax_fardal_chain_workspace_bytes(M)), 8) // 8,), dtype=torch.int64, device="cuda")
    rc = L.gx_jax_fardal_chain(int(key_data[0]), int(key_data[1]), M, out.data_ptr(), ws.data_ptr(),
                               torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "gx_jax_fardal_chain")
    return out


@dataclasses.dataclass(frozen=True)
class Fardal2015DF:
    """experimental/df.py:87-175: Fardal+15 release with adjustable means and dispersions."""

    kr_bar: float = 2.0
    kvphi_bar: float = 0.3
    kz_bar: float = 0.0
    kvz_bar: float = 0.0
    sigma_kr: float = 0.5
    sigma_kvphi: float = 0.5
    sigma_kz: float = 0.5
    sigma_kvz: float = 0.5

    def _is_default(self) -> bool:
        return self == Fardal2015DF()

    def sample(self, key, pot, /, t, x, v, Msat):
        """-> (x_lead, v_lead, x_trail, v_trail) for positions ``x`` / velocities ``v`` of shape ``(3,)`` or ``(M, 3)``.

        ``key``: an int seed or raw ``uint32[2]`` key data (reproduces ``jr.key(seed)``: pinned on the reference's
        doctest, experimental/df.py:110-123), a ``numpy.random.Generator``, or the standard-normal draws ``(4, M)``.
        The release kernel hard-wires the
        reference's legacy constants (kr = 2 + 0.5 n1, ...); other means/sigmas are mapped onto it exactly by an
        affine change of the draws (k = mean + sigma n = 2 + 0.5 n'  with  n' = (mean - 2 + sigma n) / 0.5, etc.).
        """
        torch = _lib.require_cuda()
        pot = _as_potential(pot)
        dq, restore = _to_device(x)
        dp, _ = _to_device(v)
        scalar = dq.ndim == 1
        dq, dp = dq.reshape(-1, 3).contiguous(), dp.reshape(-1, 3).contiguous()
        M = dq.shape[0]
        n = _fardal_normals(key, M).to(dq.device)
        # affine map of the draws onto the kernel's fixed constants (2, 0.3, 0, 0; all sigmas 0.5)
        dd = torch.stack([(self.kr_bar - 2.0 + self.sigma_kr * n[0]) / 0.5,
                          (self.kvphi_bar - 0.3 + self.sigma_kvphi * n[1]) / 0.5,
                          (self.kz_bar + self.sigma_kz * n[2]) / 0.5,
                          (self.kvz_bar + self.sigma_kvz * n[3]) / 0.5]).contiguous()
        mass = np.broadcast_to(np.asarray(Msat, dtype=np.float64), (M,)).copy()
        dm = torch.from_numpy(mass).to(dq.device)
        outs = [torch.empty((M, 3), dtype=torch.float64, device=dq.device) for _ in range(4)]
        import ctypes as C

        P = pot.c_struct()
        with torch.cuda.device(dq.device):
            tt = torch.from_numpy(np.broadcast_to(np.asarray(t, dtype=np.float64), (M,)).copy()).to(dq.device)
            rc = _lib.lib().gx_stream_release_t(C.byref(P), _lib.DF_FARDAL15, dq.data_ptr(), dp.data_ptr(), dm.data_ptr(),
                                                tt.data_ptr(), dd.data_ptr(), M, *[o.data_ptr() for o in outs],
                                                torch.cuda.current_stream().cuda_stream)  # fmt: skip
        _lib.check(rc, "gx_stream_release_t")
        ql, pl, qt, pt = (restore(o[0] if scalar else o) for o in outs)
        return ql, pl, qt, pt


@dataclasses.dataclass
class StreamICs:
    """experimental/stream.py:54-72."""

    release_times: Any
    prog_mass: Any
    qp_lead: tuple
    qp_trail: tuple


class StreamSimulator:
    """experimental/stream.py:75-446."""

    def init(self, pot, prog_w0, prog_t0, /, release_times, Msat, kinematic_df: Fardal2015DF | None = None, *, key,
             solver: DiffEqSolver = default_stream_solver, solver_kwargs: Mapping[str, Any] | None = None) -> StreamICs:  # fmt: skip
        """stream.py:122-237: sort the release times; integrate the progenitor from ``prog_t0`` to the first one; then
        integrate over the release times saving at each; then one DF sample per release time with the key chain
        ``key, subkey = jr.split(key)`` of the reference's ``lax.scan``."""
        from . import jaxrandom

        df = Fardal2015DF() if kinematic_df is None else kinematic_df
        rel = np.sort(np.asarray(release_times, dtype=np.float64))
        M = rel.shape[0]
        kw = dict(solver_kwargs or {})
        ctrl = kw.get("stepsize_controller", solver.stepsize_controller)
        ms = kw.get("max_steps", solver.max_steps)
        q0n = np.asarray(_host(prog_w0[0]), dtype=np.float64)[None]
        p0n = np.asarray(_host(prog_w0[1]), dtype=np.float64)[None]
        common = dict(solver=solver.solver, controller=ctrl, dt0=kw.get("dt0"), max_steps=ms)
        qa, pa, _, _ = _integrate(pot, q0n, p0n, float(prog_t0), float(rel[0]), [float(rel[0])], **common)
        xq, xp, _, _ = _integrate(pot, qa[:, 0], pa[:, 0], float(rel[0]), float(rel[-1]), rel, **common)
        xq, xp = xq[0], xp[0]
        mass = np.broadcast_to(np.asarray(Msat, dtype=np.float64), rel.shape).copy()
        if isinstance(key, (int, np.integer)) or _is_key_data(key):
            k = jaxrandom.key(key) if isinstance(key, (int, np.integer)) else key
            draws = _fardal_chain_normals(k, M)
        else:
            draws = _fardal_normals(key, M)
        ql, pl, qt, pt = df.sample(draws, pot, rel, xq, xp, mass)
        return StreamICs(release_times=rel, prog_mass=mass, qp_lead=(ql, pl), qp_trail=(qt, pt))

    def run(self, pot, stream_ics: StreamICs, /, t1, *, solver: DiffEqSolver = default_stream_solver,
            solver_kwargs: Mapping[str, Any] | None = None, throw: bool = True):  # fmt: skip
        """Integrate every released particle from its release time to ``t1`` (stream.py:239-446).
        -> ((q_lead, p_lead), (q_trail, p_trail)), each ``(M, 3)``."""
        kw = dict(solver_kwargs or {})
        ctrl = kw.get("stepsize_controller", solver.stepsize_controller)
        ms = kw.get("max_steps", solver.max_steps)
        rel = np.asarray(stream_ics.release_times, dtype=np.float64)
        M = rel.shape[0]
        q_all = _cat0(stream_ics.qp_lead[0], stream_ics.qp_trail[0])
        p_all = _cat0(stream_ics.qp_lead[1], stream_ics.qp_trail[1])
        q, p, _, _ = _integrate(pot, q_all, p_all, np.concatenate([rel, rel]), float(t1), [float(t1)],
                                solver=solver.solver, controller=ctrl, dt0=kw.get("dt0"), max_steps=ms, throw=throw)
        return (q[:M, 0], p[:M, 0]), (q[M:, 0], p[M:, 0])


def _host(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else x


__all__ = ["integrate_orbit", "Fardal2015DF", "StreamSimulator", "StreamICs", "DiffEqSolver", "default_solver",
           "default_stream_solver", "NoLoop", "Scan", "VMap", "Vectorize", "Determine", "parse_t0_t1_saveat"]  # fmt: skip
