"""CPU: the force tables of the integrators (gx_force_table, host only): the fitted polynomials, evaluated in plain
numpy exactly as the kernels do (index and argument from the bits of s, Horner), against mpmath."""
import ctypes as C

import mpmath as mp
import numpy as np
import pytest

from galax_b200 import _lib


def table(which, a=0.0):
    L = _lib.lib()
    n, deg, lo, sb, err = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_double()
    assert L.gx_force_table(which, a, None, 0, C.byref(n), C.byref(deg), C.byref(lo), C.byref(sb), None) == 0
    coef = np.empty((n.value, deg.value + 1))
    rc = L.gx_force_table(which, a, coef.ctypes.data, coef.size, None, None, None, None, C.byref(err))
    assert rc == 0
    return coef, lo.value, sb.value, err.value


def evaluate(coef, e_lo, sub_bits, s):
    """What poly_table_eval does on the device (gx_potential.cuh), in numpy."""
    s = np.asarray(s, dtype=np.float64)
    hi = (s.view(np.uint64) >> np.uint64(32)).astype(np.int64)
    j = (hi >> (20 - sub_bits)) - (1023 + e_lo) * (1 << sub_bits)
    assert (j >= 0).all() and (j < coef.shape[0]).all()
    m, e = np.frexp(s)                      # s = m 2^e, m in [0.5, 1)
    m = 2.0 * m                             # [1, 2)
    sub = (hi >> (20 - sub_bits)) & ((1 << sub_bits) - 1)
    t = m * float(2 << sub_bits) - (float(2 << sub_bits) + 2.0 * sub + 1.0)  # exact in float64
    v = coef[j, -1]
    for k in range(coef.shape[1] - 2, -1, -1):
        v = v * t + coef[j, k]
    return v


def test_nfw_force_table_on_the_host():
    coef, e_lo, sb, err = table(0)
    assert coef.shape == (384, 10) and e_lo == -7 and sb == 5 and err < 3e-16
    rng = np.random.default_rng(3)
    edges = np.ldexp(1.0 + np.arange(32) / 32.0, rng.integers(-7, 5, 32))
    below = np.nextafter(edges, 0)
    s = np.concatenate([2.0 ** rng.uniform(-7, 5, 400), edges, below[below >= 2.0**-7], [2.0**-7, np.nextafter(32.0, 0)]])
    F = evaluate(coef, e_lo, sb, s)
    mp.mp.dps = 30
    ref = np.array([float((mp.log1p(mp.mpf(float(x))) - mp.mpf(float(x)) / (1 + mp.mpf(float(x)))) / mp.mpf(float(x)) ** 3) for x in s])
    assert np.abs(F / ref - 1).max() < 6e-16  # (numpy's Horner is not fused: a little above the device's 5e-16)


@pytest.mark.parametrize("a", [0.6, 1.05, 0.25])
def test_powerlawcutoff_force_table_on_the_host(a):
    coef, e_lo, sb, err = table(1, a)
    assert coef.shape == (448, 10) and e_lo == -11 and sb == 5 and err < 1e-15
    rng = np.random.default_rng(4)
    s = 2.0 ** rng.uniform(-11, 3, 200)
    G = evaluate(coef, e_lo, sb, s)
    mp.mp.dps = 30
    ref = np.array([float(mp.gammainc(a, 0, mp.mpf(float(x)) ** 2, regularized=True) / mp.mpf(float(x)) ** 3) for x in s])
    assert np.abs(G / ref - 1).max() < 1e-15


def test_bad_arguments():
    L = _lib.lib()
    assert L.gx_force_table(2, 0.6, None, 0, None, None, None, None, None) == -1
    assert L.gx_force_table(1, -1.0, None, 0, None, None, None, None, None) == -1
    buf = np.empty(10)
    assert L.gx_force_table(0, 0.0, buf.ctypes.data, 10, None, None, None, None, None) == -1  # capacity too small
