"""ctypes binding of ``oracle/libgalax_oracle.so`` (TEST INFRASTRUCTURE; see ``oracle/__init__.py``)."""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

from . import dopri8_tableau as tabmod
from . import potentials as op

_HERE = Path(__file__).resolve().parent
_LIB = None

OK, MAX_STEPS, NONFINITE = 0, 1, 2


class _Component(C.Structure):
    _fields_ = [("kind", C.c_int), ("group", C.c_int), ("p", C.c_double * 8), ("dp", C.c_double * 8)]


class _Potential(C.Structure):
    _fields_ = [("n", C.c_int), ("G", C.c_double), ("c", _Component * 16)]


class _Tableau(C.Structure):
    _fields_ = [
        ("ns", C.c_int),
        ("order", C.c_double),
        ("a", (C.c_double * 14) * 14),
        ("b_sol", C.c_double * 14),
        ("b_err", C.c_double * 14),
        ("c", C.c_double * 14),
        ("dense", (C.c_double * 6) * 14),
    ]


class _Pid(C.Structure):
    _fields_ = [
        ("rtol", C.c_double), ("atol", C.c_double),
        ("pcoeff", C.c_double), ("icoeff", C.c_double), ("dcoeff", C.c_double),
        ("safety", C.c_double), ("factormin", C.c_double), ("factormax", C.c_double),
        ("dtmin", C.c_double), ("dtmax", C.c_double),
        ("force_dtmin", C.c_int),
        ("dt0", C.c_double),
    ]  # fmt: skip


def build(force: bool = False) -> Path:
    so = _HERE / "libgalax_oracle.so"
    src = _HERE / "galax_oracle.c"
    hdr = _HERE.parent / "include" / "gx_portable_math.h"
    if force or not so.exists() or so.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(str(build()))
        _LIB.oc_potential_value.restype = C.c_double
        _LIB.oc_gammainc.restype = C.c_double
        _LIB.oc_gammainc.argtypes = [C.c_double, C.c_double]
        _LIB.oc_num_threads.restype = C.c_int
    return _LIB


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def c_potential(pot: op.Potential) -> _Potential:
    P = _Potential()
    P.n = len(pot.components)
    P.G = pot.G
    gid = {}
    for g, idxs in enumerate(pot.group_list()):
        for i in idxs:
            gid[i] = g
    for i, comp in enumerate(pot.components):
        P.c[i].kind = comp.kind
        P.c[i].group = gid[i]
        for j, v in enumerate(comp.params):
            P.c[i].p[j] = v
        for j, v in enumerate(comp.rates):
            P.c[i].dp[j] = v
    return P


def c_tableau(solver: str = "dopri8") -> _Tableau:
    from . import dopri5_tableau as tab5

    mod, ns, order = (tabmod, 14, 8.0) if solver == "dopri8" else (tab5, 7, 5.0)
    t = _Tableau()
    t.ns, t.order = ns, order
    A = mod.a_matrix()
    for i in range(ns):
        for j in range(ns):
            t.a[i][j] = A[i, j]
    for name, vec in (("b_sol", mod.b_sol()), ("b_err", mod.b_err()), ("c", mod.c_vec())):
        arr = getattr(t, name)
        for i in range(ns):
            arr[i] = vec[i]
    Bd = mod.dense_b()
    for i in range(ns):
        for m in range(6):
            t.dense[i][m] = Bd[i, m]
    return t


def num_threads() -> int:
    return int(lib().oc_num_threads())


def use_all_cores() -> int:
    """Override OMP_NUM_THREADS (torchrun sets it to 1): use every core this process may run on."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().oc_set_num_threads(C.c_int(n))
    return num_threads()


def gammainc(a: float, x: float) -> float:
    return float(lib().oc_gammainc(a, x))


def potential_eval(pot: op.Potential, xyz, what=("phi", "grad", "acc", "hess"), t: float = 0.0):
    xyz = _f64(xyz).reshape(-1, 3)
    N = xyz.shape[0]
    P = c_potential(pot)
    mask = sum({"phi": 1, "grad": 2, "acc": 4, "hess": 8}[w] for w in what)
    phi = np.empty(N)
    grad = np.empty((N, 3))
    acc = np.empty((N, 3))
    hess = np.empty((N, 3, 3))
    lib().oc_potential_eval_t(C.byref(P), C.c_double(t), C.c_int64(N), _dp(xyz), C.c_uint(mask), _dp(phi), _dp(grad),
                              _dp(acc), _dp(hess))
    out = {"phi": phi, "grad": grad, "acc": acc, "hess": hess}
    return {k: out[k] for k in what}


def integrate_fixed(pot, q0, p0, t0, t1, dt0, ts, *, scheme=0, max_steps=-1):
    """SemiImplicitEuler (scheme 0) or LeapfrogMidpoint (scheme 1) with ConstantStepSize.

    Returns q[N,T,3], p[N,T,3], status[N], nsteps[N].
    """
    q0, p0, ts = _f64(q0).reshape(-1, 3), _f64(p0).reshape(-1, 3), _f64(ts).reshape(-1)
    N, T = q0.shape[0], ts.shape[0]
    q = np.empty((N, T, 3))
    p = np.empty((N, T, 3))
    status = np.zeros(N, dtype=np.int32)
    nsteps = np.zeros(N, dtype=np.int64)
    P = c_potential(pot)
    rc = lib().oc_integrate_fixed(
        C.byref(P), C.c_int64(N), _dp(q0), _dp(p0), C.c_double(t0), C.c_double(t1), C.c_double(dt0),
        C.c_int(T), _dp(ts), C.c_int(scheme), C.c_int64(-1 if max_steps is None else max_steps),
        _dp(q), _dp(p), _ip(status), nsteps.ctypes.data_as(C.POINTER(C.c_int64)),
    )  # fmt: skip
    if rc != 0:
        raise ValueError("oc_integrate_fixed: bad arguments")
    return q, p, status, nsteps


def make_pid(rtol, atol, *, pcoeff=0.0, icoeff=1.0, dcoeff=0.0, safety=0.9, factormin=0.2, factormax=10.0,
             dtmin=None, dtmax=None, force_dtmin=True, dt0=None) -> _Pid:  # fmt: skip
    pid = _Pid()
    pid.rtol, pid.atol = rtol, atol
    pid.pcoeff, pid.icoeff, pid.dcoeff = pcoeff, icoeff, dcoeff
    pid.safety, pid.factormin, pid.factormax = safety, factormin, factormax
    pid.dtmin = -1.0 if dtmin is None else dtmin
    pid.dtmax = -1.0 if dtmax is None else dtmax
    pid.force_dtmin = int(force_dtmin)
    pid.dt0 = -1.0 if dt0 is None else dt0
    return pid


def integrate_dopri8(pot, q0, p0, t0, t1, ts, *, rtol=1e-8, atol=1e-8, max_steps=-1, solver="dopri8", **pid_kw):
    """Per-particle Dopri8 (or ``solver="dopri5"``) + PID.  ``t0`` scalar or array[N].
    Returns q, p, status, n_accepted, n_attempted."""
    q0, p0, ts = _f64(q0).reshape(-1, 3), _f64(p0).reshape(-1, 3), _f64(ts).reshape(-1)
    N, T = q0.shape[0], ts.shape[0]
    t0a = _f64(t0).reshape(-1)
    stride = 0 if t0a.shape[0] == 1 else 1
    if stride:
        assert t0a.shape[0] == N
    q = np.empty((N, T, 3))
    p = np.empty((N, T, 3))
    status = np.zeros(N, dtype=np.int32)
    nacc = np.zeros(N, dtype=np.int32)
    ntot = np.zeros(N, dtype=np.int32)
    P, tab, pid = c_potential(pot), c_tableau(solver), make_pid(rtol, atol, **pid_kw)
    lib().oc_integrate_dopri8(
        C.byref(P), C.byref(tab), C.byref(pid), C.c_int64(N), _dp(q0), _dp(p0), _dp(t0a), C.c_int(stride),
        C.c_double(t1), C.c_int(T), _dp(ts), C.c_int64(-1 if max_steps is None else max_steps),
        _dp(q), _dp(p), _ip(status), _ip(nacc), _ip(ntot),
    )  # fmt: skip
    return q, p, status, nacc, ntot


def release_fardal(pot, xq, xp, mass, normals):
    xq, xp = _f64(xq).reshape(-1, 3), _f64(xp).reshape(-1, 3)
    M = xq.shape[0]
    mass = _f64(np.broadcast_to(np.asarray(mass, dtype=np.float64), (M,)))
    normals = _f64(normals).reshape(4, M)
    outs = [np.empty((M, 3)) for _ in range(4)]
    P = c_potential(pot)
    lib().oc_release_fardal(C.byref(P), C.c_int64(M), _dp(xq), _dp(xp), _dp(mass), _dp(normals), *[_dp(o) for o in outs])
    return tuple(outs)  # q_lead, p_lead, q_trail, p_trail


def release_chen(pot, xq, xp, mass, posvel):
    xq, xp = _f64(xq).reshape(-1, 3), _f64(xp).reshape(-1, 3)
    M = xq.shape[0]
    mass = _f64(np.broadcast_to(np.asarray(mass, dtype=np.float64), (M,)))
    posvel = _f64(posvel).reshape(M, 6)
    outs = [np.empty((M, 3)) for _ in range(4)]
    P = c_potential(pot)
    lib().oc_release_chen(C.byref(P), C.c_int64(M), _dp(xq), _dp(xp), _dp(mass), _dp(posvel), *[_dp(o) for o in outs])
    return tuple(outs)


def mockstream(pot, prog_q0, prog_p0, ts, prog_mass, draws, *, df="fardal", rtol=1e-7, atol=1e-7, max_steps=-1):
    """MockStreamGenerator.run restated (mockstream_generator.py:160-275) given the random draws.

    progenitor orbit saved at ``ts`` (one adaptive solve) -> DF release -> each particle integrated
    from ts[i] to t_f = ts[-1] + 1e-3, final state kept.
    Returns dict(lead_q, lead_p, trail_q, trail_p, prog_q, prog_p).
    """
    ts = _f64(ts)
    if ts[1] < ts[0]:
        ts = ts[::-1].copy()
    pq, pp, st, _, _ = integrate_dopri8(pot, prog_q0, prog_p0, ts[0], ts[-1], ts, rtol=rtol, atol=atol, max_steps=max_steps)
    assert st[0] == OK
    pq, pp = pq[0], pp[0]
    rel = release_fardal if df == "fardal" else release_chen
    ql, pl, qt, pt = rel(pot, pq, pp, prog_mass, draws)
    t_f = ts[-1] + 1e-3
    out = {}
    for name, (q, p) in (("lead", (ql, pl)), ("trail", (qt, pt))):
        qf, pf, st, _, _ = integrate_dopri8(pot, q, p, ts, t_f, [t_f], rtol=rtol, atol=atol, max_steps=max_steps)
        assert (st == OK).all()
        out[name + "_q"], out[name + "_p"] = qf[:, 0], pf[:, 0]
    out["prog_q"], out["prog_p"] = pq, pp
    out["release_q"] = {"lead": ql, "trail": qt}
    out["release_p"] = {"lead": pl, "trail": pt}
    return out
