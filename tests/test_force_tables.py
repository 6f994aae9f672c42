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


# ---- the combined spherical table S(r^2) of a composite (gx_spherical_force_table) ------------------------------
def sph_table(pot):
    L = _lib.lib()
    fn = L.gx_spherical_force_table
    cs = pot.c_struct()
    n, deg, lo, sb, err = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_double()
    assert fn(C.byref(cs), None, 0, C.byref(n), C.byref(deg), C.byref(lo), C.byref(sb), None) == 0
    coef = np.empty((n.value, deg.value + 1))
    assert fn(C.byref(cs), coef.ctypes.data, coef.size, None, None, None, None, C.byref(err)) == 0
    return coef, lo.value, sb.value, err.value


def evaluate_estrin(coef, e_lo, sub_bits, s):
    """The Estrin form the Dopri kernels use (sph_wide_eval<ESTRIN = true>: degree 5), in numpy."""
    s = np.asarray(s, dtype=np.float64)
    hi = (s.view(np.uint64) >> np.uint64(32)).astype(np.int64)
    j = (hi >> (20 - sub_bits)) - (1023 + e_lo) * (1 << sub_bits)
    m = 2.0 * np.frexp(s)[0]
    sub = (hi >> (20 - sub_bits)) & ((1 << sub_bits) - 1)
    t = m * float(2 << sub_bits) - (float(2 << sub_bits) + 2.0 * sub + 1.0)
    c = coef[j]
    assert c.shape[1] == 6
    t2 = t * t
    p = [c[:, 2 * k + 1] * t + c[:, 2 * k] for k in range(3)]
    return (p[2] * t2 + p[1]) * t2 + p[0]


def sph_reference(pot, r):
    """sum of Phi'(r)/r over the spherical components, 30 digits (closed forms of the reference's potentials:
    hernquist.py:60-66, nfw/base.py:326-338, powerlawcutoff.py:88-117)."""
    import galax_b200._lib as L_
    mp.mp.dps = 30
    cs = pot.c_struct()
    out = []
    for x in r:
        x = mp.mpf(float(x)); tot = mp.mpf(0)
        for i in range(cs.n):
            c = cs.c[i]
            GM = mp.mpf(float(np.float64(cs.G) * np.float64(c.p[0])))  # G m as the kernels form it (one fp64 product)
            if c.kind == L_.KIND_HERNQUIST:
                tot += GM / (x * (x + mp.mpf(c.p[1])) ** 2)
            elif c.kind == L_.KIND_NFW:
                s = x / mp.mpf(c.p[1])
                tot += GM * (mp.log1p(s) - s / (1 + s)) / x ** 3
            elif c.kind == L_.KIND_PLC:
                a = mp.mpf(1.5) - mp.mpf(c.p[1]) / 2
                tot += GM * mp.gammainc(a, 0, (x / mp.mpf(c.p[2])) ** 2, regularized=True) / x ** 3
        out.append(float(tot))
    return np.array(out)


@pytest.mark.parametrize("name", ["MilkyWayPotential", "MilkyWayPotential2022", "BovyMWPotential2014"])
def test_combined_spherical_table_on_the_host(name):
    """What the integrators look up: 128 intervals per octave of r^2, degree 5, 48-byte rows; Horner (fixed step) and
    Estrin (Dopri) evaluation, against the 30-digit closed forms."""
    import galax_b200.potential as gp

    pot = getattr(gp, name)()
    coef, e_lo, sb, err = sph_table(pot)
    assert coef.shape == (2816, 6) and e_lo == -8 and sb == 7 and err < 4e-16
    rng = np.random.default_rng(7)
    edges = np.ldexp(1.0 + np.arange(0, 128, 5) / 128.0, rng.integers(-8, 14, 26))
    below = np.nextafter(edges, 0)
    u = np.concatenate([2.0 ** rng.uniform(-8, 14, 300), edges, below[below >= 2.0**-8], [2.0**-8, np.nextafter(2.0**14, 0)]])
    ref = sph_reference(pot, [mp.sqrt(mp.mpf(float(x))) for x in u])
    for ev in (evaluate, evaluate_estrin):
        assert np.abs(ev(coef, e_lo, sb, u) / ref - 1).max() < 8e-16, ev.__name__


def test_combined_spherical_table_follows_the_scale_radii():
    """The 22 octaves end at (8 x the largest scale radius)^2 rounded up to a power of two: the library does not know the
    unit system, the scale radii do.  Any composite of Hernquist / NFW / PowerLawCutoff terms gets one."""
    import galax_b200.potential as gp

    cases = [
        (gp.CompositePotential(d=gp.MiyamotoNagaiPotential(5e10, 3.0, 0.3), b=gp.HernquistPotential(4e9, 0.5), h=gp.NFWPotential(8e11, 20.0)), -6),
        (gp.CompositePotential(b=gp.HernquistPotential(1e6, 2e-3), h=gp.NFWPotential(1e8, 0.1)), -22),     # a dwarf, in kpc
        (gp.CompositePotential(b=gp.HernquistPotential(4e9, 500.0), h=gp.NFWPotential(8e11, 2e4)), 14),      # the same galaxy in pc
        (gp.KeplerPotential(1e12), -8),                                                                       # no scale: default
    ]
    rng = np.random.default_rng(6)
    for pot, want in cases:
        coef, e_lo, sb, err = sph_table(pot)
        assert e_lo == want and coef.shape == (2816, 6) and err < 6e-16, (want, e_lo, err)
        u = 2.0 ** rng.uniform(e_lo, e_lo + 22, 120)
        ref = sph_reference(pot, [mp.sqrt(mp.mpf(float(x))) for x in u])
        assert np.abs(evaluate(coef, e_lo, sb, u) / ref - 1).max() < 8e-16


def test_combined_spherical_table_bad_arguments():
    import galax_b200.potential as gp

    L = _lib.lib()
    cs = gp.MiyamotoNagaiPotential(m_tot=1e10, a=3.0, b=0.3).c_struct()
    assert L.gx_spherical_force_table(C.byref(cs), None, 0, None, None, None, None, None) == -2  # GX_ERR_UNSUPPORTED: no spherical component
    cs = gp.MilkyWayPotential().c_struct()
    buf = np.empty(10)
    assert L.gx_spherical_force_table(C.byref(cs), buf.ctypes.data, 10, None, None, None, None, None) == -1
