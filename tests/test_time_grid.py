"""CPU: the run-length form of diffrax's ConstantStepSize time grid (gx_fixed_time_grid, host only -- no CUDA call).

The fixed-step kernel of the three named Milky-Way models consumes the grid as runs of equal, exactly representable
steps; these tests check the runs against a plain walk of the recurrence t_{n+1} = fl(t_n + dt0) with diffrax's
``_clip_to_end`` (tolerance 1e-10), the same recurrence the oracle integrates on (oracle/galax_oracle.c)."""
import ctypes as C
from fractions import Fraction

import numpy as np
import pytest

from galax_b200 import _lib


def walk(t0, t1, dt0, max_steps=None):
    """The grid times in tau = dir * t, as a list starting at dir * t0."""
    d = 1.0 if t1 >= t0 else -1.0
    T0, T1, h0 = t0 * d, t1 * d, dt0 * d
    clip = lambda tn: T1 if tn > T1 - 1e-10 else tn
    ts, tprev, tnext = [T0], T0, clip(T0 + h0)
    while tprev < T1:
        if max_steps is not None and len(ts) - 1 >= max_steps:
            return ts, True
        ts.append(tnext)
        tprev, tnext = tnext, clip(tnext + h0)
    return ts, False


def runs(t0, t1, dt0, max_steps=-1, cap=128):
    L = _lib.lib()
    n, hit, nr = C.c_int64(), C.c_int32(), C.c_int32()
    cnt, h = (C.c_int64 * cap)(), (C.c_double * cap)()
    rc = L.gx_fixed_time_grid(t0, t1, dt0, max_steps, C.byref(n), C.byref(hit), C.byref(nr), cnt, h, cap)
    assert rc == 0
    k = max(nr.value, 0)
    return n.value, bool(hit.value), nr.value, list(cnt[:k]), list(h[:k])


CASES = [
    (0.0, 1000.0, 0.1),        # C1 / the bench
    (0.0, 100.0, 0.1), (0.0, 37.3, 0.25), (-40.0, 55.5, 0.07), (30.0, -20.0, -0.13), (1000.0, 1003.0, 1e-3),
    (0.0, 10.0, 3.0), (5.0, 5.0, 0.1), (0.0, 1.0, 2.0), (0.0, 200.0, 0.001), (-3000.0, 0.0, 0.5), (1e-3, 2e-3, 1e-7),
    (0.0, 1.0, 1.0 / 3.0), (123.456, 789.0123, 0.0317),
]


@pytest.mark.parametrize("t0,t1,dt0", CASES)
def test_runs_reproduce_the_grid_exactly(t0, t1, dt0):
    ts, _ = walk(t0, t1, dt0)
    n, hit, nr, cnt, h = runs(t0, t1, dt0)
    assert n == len(ts) - 1 and not hit
    assert nr >= 0 and sum(cnt) == n and all(c > 0 for c in cnt)
    # inside a run the grid times are t_s + j h exactly (what the kernel's save-time search relies on), and the runs
    # chain: the end of one is the start of the next
    t, i = ts[0], 0
    for c, hh in zip(cnt, h):
        j = np.arange(1, c + 1, dtype=np.float64)
        # fma(j, h, t_s) on the device: one rounding of the exact t_s + j h (here through Fraction)
        got = np.array([float(Fraction(t) + int(jj) * Fraction(hh)) for jj in j[: min(c, 2000)]])  # = fma(j, h, t_s)
        assert np.array_equal(got, np.array(ts[i + 1 : i + 1 + len(got)])), (t0, t1, dt0)
        assert ts[i + c] - ts[i + c - 1] == hh
        t, i = ts[i + c], i + c
    assert i == n


def test_c1_grid_is_a_few_dozen_runs_and_counts_follow_max_steps():
    n, hit, nr, cnt, h = runs(0.0, 1000.0, 0.1)
    assert n == 10_000 and 10 <= nr <= 60
    assert abs(sum(c * hh for c, hh in zip(cnt, h)) - 1000.0) < 1e-9
    n, hit, nr, cnt, h = runs(0.0, 1000.0, 0.1, max_steps=600)
    assert n == 600 and hit and sum(cnt) == 600
    n, hit, nr, cnt, h = runs(5.0, 5.0, 0.1)
    assert n == 0 and nr == 0 and not hit


def test_random_grids_and_fallback():
    rng = np.random.default_rng(7)
    for _ in range(60):
        t0 = float(rng.uniform(-500, 500))
        span = float(10 ** rng.uniform(-2, 3)) * (1 if rng.random() < 0.7 else -1)
        dt0 = abs(span) / float(rng.integers(3, 4000)) * (1 if span > 0 else -1)
        ts, _ = walk(t0, t0 + span, dt0)
        n, hit, nr, cnt, h = runs(t0, t0 + span, dt0)
        assert n == len(ts) - 1
        if nr >= 0:
            assert sum(cnt) == n
            steps = np.diff(np.array(ts))
            assert np.array_equal(np.repeat(np.array(h), np.array(cnt, dtype=np.int64)), steps)
    # wrong direction of dt0 is an argument error; a grid needing more runs than the kernel takes reports -1
    L = _lib.lib()
    assert L.gx_fixed_time_grid(0.0, 1.0, -0.1, -1, None, None, None, None, None, 0) == -1  # GX_ERR_BADARG
    # every sum a tie (dt0 an odd multiple of half an ulp of t): round-to-even settles on one step size after the first
    # step, so even this grid is two runs
    dt0 = (2.0**40 + 1) * 2.0**-53
    ts, _ = walk(1.0, 1.1, dt0)
    n, hit, nr, cnt, h = runs(1.0, 1.1, dt0)
    assert n == len(ts) - 1 and 1 <= nr <= 3
    assert np.array_equal(np.repeat(np.array(h), np.array(cnt, dtype=np.int64)), np.diff(np.array(ts)))


def test_pipeline_gate():
    """dynamics._pipeline_ok: large contiguous float64 host batches (numpy or torch) take the chunked copy/compute
    pipeline; everything else is one launch (decided without touching a device)."""
    import torch

    import galax_b200.dynamics as gd

    n = gd.PIPELINE_MIN_PARTICLES
    q = torch.zeros((n, 3), dtype=torch.float64)
    ts = np.array([1.0])
    assert gd._pipeline_ok(torch, q.numpy(), q.numpy(), 0.0, ts, "NT3")          # large numpy batch: pipelined
    assert gd._pipeline_ok(torch, q, q, 0.0, ts, "NT3")                          # ... and host tensors, pinned or not
    assert not gd._pipeline_ok(torch, q.numpy(), q.numpy()[::-1], 0.0, ts, "NT3")  # not contiguous
    assert not gd._pipeline_ok(torch, q.numpy(), q, 0.0, ts, "NT3")              # mixed kinds
    assert not gd._pipeline_ok(torch, q, q, 0.0, np.zeros(40_000), "NT3")        # result too large for pinned buffers
    assert not gd._pipeline_ok(torch, q[:10], q[:10], 0.0, ts, "NT3")            # small
    assert not gd._pipeline_ok(torch, q, q, 0.0, ts, "T3N")                      # layout that cannot be sliced by particle
    assert not gd._pipeline_ok(torch, q.float(), q.float(), 0.0, ts, "NT3")      # dtype


def test_trip_count_agrees_with_the_oracle_integrator():
    """The library's host walk and the C oracle's integrator (oracle/galax_oracle.c) take the same number of steps on
    the same grids, max_steps included."""
    from oracle import cref
    from oracle import potentials as op

    pot = op.milky_way_potential()
    q0, p0 = np.array([[8.0, 0.0, 0.5]]), np.array([[0.0, 0.2, 0.01]])
    for t0, t1, dt0 in CASES:
        if t0 == t1 or abs((t1 - t0) / dt0) > 30_000:
            continue
        _, _, st, n = cref.integrate_fixed(pot, q0, p0, t0, t1, dt0, [t1])
        n_lib, hit, *_ = runs(t0, t1, dt0)
        assert int(n[0]) == n_lib and not hit and st[0] == 0, (t0, t1, dt0)
    _, _, st, n = cref.integrate_fixed(pot, q0, p0, 0.0, 100.0, 0.1, [100.0], max_steps=77)
    n_lib, hit, *_ = runs(0.0, 100.0, 0.1, max_steps=77)
    assert int(n[0]) == n_lib == 77 and hit and st[0] == 1
