"""TEST INFRASTRUCTURE ONLY.  numpy restatement of ``diffrax.diffeqsolve(Dopri8, PIDController)`` on the JOINT state.

The reference's scalar-time batch path (``OrbitSolver.solve(field, (q[N,3], p[N,3]), t0, t1)`` with scalar times,
``dynamics/_src/orbit/solver.py:774-803``; ``OrbitSolver.run``, ``:358-380``) hands the whole ``(N, 3)`` batch to
one ``diffeqsolve``: a single 6N-dimensional ODE, ONE step size, the error norm an RMS over all 6N numbers.  The
product path (and ``galax_oracle.c``) controls the step per particle instead (what the reference does under
``vmap`` / batched times), so this module exists to pin the restated algorithm -- tableau,
controller, initial step -- on the doctests of that joint path, and as an independent (generic first-order form,
no Nystrom rewriting) cross-check of the C oracle for N = 1.

Generic form: k_i = f(y0 + h sum_j a_ij k_j); y1 = y0 + h sum b_i k_i; y_err = h sum (b - b_hat)_i k_i.
"""
from __future__ import annotations

import numpy as np

from . import dopri8_tableau as tab
from . import potentials as op


def _rms(x: np.ndarray) -> float:
    return float(np.sqrt(np.mean(np.square(x))))


def select_initial_step(f, t0, y0, rtol, atol, error_order, *, exponent_order=None):
    """diffrax ``PIDController._select_initial_step`` (Hairer, Norsett & Wanner II.4)."""
    k = error_order + 1.0 if exponent_order is None else exponent_order
    scale = atol + np.abs(y0) * rtol
    f0 = f(t0, y0)
    d0, d1 = _rms(y0 / scale), _rms(f0 / scale)
    cond = d0 < 1e-5 or d1 < 1e-5
    h0 = 1e-6 if cond else 0.01 * (d0 / d1)
    f1 = f(t0 + h0, y0 + h0 * f0)
    d2 = _rms((f1 - f0) / scale) / h0
    max_d = max(d1, d2)
    h1 = max(1e-6, h0 * 1e-3) if max_d <= 1e-15 else (0.01 / max_d) ** (1.0 / k)
    return min(100.0 * h0, h1)


def step(f, t, y, h, k1=None):
    """One Dopri8 step; returns (y1, y_error, stages[14, ...])."""
    A, b, be, c = tab.a_matrix(), tab.b_sol(), tab.b_err(), tab.c_vec()
    ks = np.zeros((tab.N_STAGES,) + y.shape)
    ks[0] = f(t, y) if k1 is None else k1
    for i in range(1, tab.N_STAGES):
        incr = np.tensordot(A[i, :i], ks[:i], axes=1)
        ks[i] = f(t + c[i] * h, y + h * incr)
    y1 = y + h * np.tensordot(b, ks, axes=1)
    yerr = h * np.tensordot(be, ks, axes=1)
    return y1, yerr, ks


def solve(pot: op.Potential, q0, p0, t0, t1, ts=None, *, rtol=1e-8, atol=1e-8, dt0=None, dtmin=None, dtmax=None,
          safety=0.9, factormin=0.2, factormax=10.0, max_steps=4096, error_order=8.0, exponent_order=None):
    """Joint solve of the (N, 3) batch with one shared adaptive step.  Returns (q[T,N,3], p[T,N,3], stats)."""
    q0 = np.atleast_2d(np.asarray(q0, float))
    p0 = np.atleast_2d(np.asarray(p0, float))
    p0 = np.broadcast_to(p0, q0.shape).copy()
    n = q0.shape[0]

    def f(t, y):
        return np.concatenate([y[n:], -op.gradient(pot, y[:n])], axis=0)

    y = np.concatenate([q0, p0], axis=0)
    ts = np.array([t1], float) if ts is None else np.atleast_1d(np.asarray(ts, float))
    out = np.full((len(ts), 2 * n, 3), np.nan)
    direction = 1.0 if t1 >= t0 else -1.0
    # diffrax integrates backwards by flipping the sign of time; do the same
    if direction < 0:
        raise NotImplementedError("forward only")
    if dt0 is None:
        dt0 = select_initial_step(f, t0, y, rtol, atol, error_order, exponent_order=exponent_order)
    if dtmax is not None:
        dt0 = min(dt0, dtmax)
    at_dtmin = False
    if dtmin is not None:
        at_dtmin = dt0 <= dtmin
        dt0 = max(dt0, dtmin)
    tprev, tnext = t0, t0 + dt0
    if tnext > t1 - 1e-10:
        tnext = t1
    out[ts == t0] = y
    k1 = None
    n_acc = n_tot = 0
    while tprev < t1 and n_tot < max_steps:
        h = tnext - tprev
        y1, yerr, ks = step(f, tprev, y, h, k1)
        n_tot += 1
        scale = atol + rtol * np.maximum(np.abs(y), np.abs(y1))
        serr = _rms(yerr / scale)
        keep = serr < 1.0
        if dtmin is not None:
            keep = keep or at_dtmin
        if not np.isfinite(serr):
            keep, factor = False, factormin
        elif serr == 0.0:
            factor = factormax
        else:
            factor = min(factormax, max(1.0 if keep else factormin, safety * serr ** (-1.0 / error_order)))
        if keep:
            n_acc += 1
            sel = (ts > tprev) & (ts <= tnext)
            for idx in np.nonzero(sel)[0]:
                th = (ts[idx] - tprev) / h
                out[idx] = y1 if ts[idx] == tnext else y + h * np.tensordot(tab.dense_weights(th), ks, axes=1)
            y, k1 = y1, ks[-1]
            tprev = tnext
        dt = h * factor
        if dtmax is not None:
            dt = min(dt, dtmax)
        if dtmin is not None:
            at_dtmin = dt <= dtmin
            dt = max(dt, dtmin)
        tnext = tprev + dt
        if tnext > t1 - 1e-10:
            tnext = t1
    return out[:, :n], out[:, n:], {"num_steps": n_tot, "num_accepted_steps": n_acc, "dt0": dt0}
