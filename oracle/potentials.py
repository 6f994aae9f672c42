"""numpy fp64 restatement of the reference's analytic potentials (TEST INFRASTRUCTURE).

Every ``potential_*`` below follows the reference's closed form line by line; the
``gradient_*`` / ``hessian_*`` functions are the hand-derived derivatives of those closed
forms (the reference obtains them with ``jax.grad`` / ``jax.hessian``,
``/root/reference/src/galax/potential/_src/base.py:170-179,230-239``).
``tests/test_oracle_potentials.py::test_hand_derivatives_vs_mpmath_differentiation`` checks the hand derivations
against 40-digit mpmath differentiation of the *potential* itself.

Unit system: galactic (kpc, Myr, Msun, rad).  All arrays are fp64; ``xyz`` has shape
``(..., 3)``.
"""

from __future__ import annotations

import dataclasses
from typing import Sequence

import numpy as np
from scipy import special as sps

# G = astropy CODATA2018 ``G`` decomposed into galactic units, as the reference does at
# ``potential/_src/base.py:35,91-95``.  kpc^3 / (Msun Myr^2).
G_GALACTIC = 4.498502151469553e-12
# smallest normal double, added under the square root by ``safe_sqrt``
# (``potential/_src/utils.py:44-79``).
TINY = float(np.finfo(np.float64).tiny)

KIND_MN = 0  # MiyamotoNagaiPotential      p = (m_tot, a, b)
KIND_HERNQUIST = 1  # HernquistPotential    p = (m_tot, r_s)
KIND_NFW = 2  # NFWPotential                p = (m, r_s)
KIND_PLC = 3  # PowerLawCutoffPotential     p = (m_tot, alpha, r_c)
KIND_LOG = 4  # (LMJ09)LogarithmicPotential p = (v_c, r_s, q1, q2, q3, phi)
KIND_ISOCHRONE = 5  # IsochronePotential     p = (m_tot, r_s)
KIND_SATOH = 6  # SatohPotential             p = (m_tot, a, b)
KIND_TRIAXIAL_HERNQUIST = 7  # TriaxialHernquistPotential p = (m_tot, r_s, q1, q2)
KIND_JAFFE = 8  # JaffePotential             p = (m_tot, r_s)
KIND_BURKERT = 9  # BurkertPotential         p = (m, r_s)
KIND_STONE = 10  # StoneOstriker15Potential  p = (m_tot, r_c, r_h)
KIND_HARMONIC = 11  # HarmonicOscillatorPotential p = (omega_x, omega_y, omega_z)
KIND_HENON_HEILES = 12  # HenonHeilesPotential p = (coeff, timescale)
KIND_NAMES = {KIND_TRIAXIAL_HERNQUIST: "TriaxialHernquist", KIND_JAFFE: "Jaffe", KIND_BURKERT: "Burkert", KIND_STONE: "StoneOstriker15",
              KIND_HARMONIC: "HarmonicOscillator", KIND_HENON_HEILES: "HenonHeiles",
              KIND_MN: "MN", KIND_HERNQUIST: "Hernquist", KIND_NFW: "NFW", KIND_PLC: "PowerLawCutoff",
              KIND_LOG: "Logarithmic", KIND_ISOCHRONE: "Isochrone", KIND_SATOH: "Satoh"}


@dataclasses.dataclass(frozen=True)
class Component:
    kind: int
    params: tuple[float, ...]
    name: str = ""
    rates: tuple[float, ...] = ()  # LinearParameter (params/core.py:25-110): p_k(t) = params[k] + rates[k] * t

    def params_at(self, t: float) -> tuple[float, ...]:
        if not self.rates:
            return self.params
        r = tuple(self.rates) + (0.0,) * (len(self.params) - len(self.rates))
        return tuple(p + dp * t for p, dp in zip(self.params, r))


@dataclasses.dataclass(frozen=True)
class Potential:
    """A composite: components are summed in list order (``base_multi.py:39-82``).

    ``groups`` records which consecutive components were one reference component (the three
    MN disks of an MN3 model are summed first, ``mn3.py:121-130``), so the oracle can
    reproduce the reference's summation tree.
    """

    components: tuple[Component, ...]
    G: float = G_GALACTIC
    groups: tuple[tuple[int, ...], ...] = ()
    name: str = "composite"

    def group_list(self) -> tuple[tuple[int, ...], ...]:
        if self.groups:
            return self.groups
        return tuple((i,) for i in range(len(self.components)))


# ----------------------------------------------------------------------------------------
# helpers


def _r_safe(xyz: np.ndarray) -> np.ndarray:
    """``r_spherical`` -> ``safe_vector_norm`` -> ``safe_sqrt`` (utils.py:44-125)."""
    xyz = np.asarray(xyz, dtype=np.float64)
    return np.sqrt(np.sum(np.square(xyz), axis=-1) + TINY)


def _spherical_grad(xyz, r, dphi_dr):
    return (dphi_dr / r)[..., None] * xyz


def _spherical_hess(xyz, r, dphi_dr, d2phi_dr2):
    n = xyz / r[..., None]
    nn = n[..., :, None] * n[..., None, :]
    eye = np.eye(3)
    return d2phi_dr2[..., None, None] * nn + (dphi_dr / r)[..., None, None] * (eye - nn)


# ----------------------------------------------------------------------------------------
# Miyamoto-Nagai (builtin/miyamotonagai.py:73-78)


def potential_mn(G, m, a, b, xyz):
    x, y, z = xyz[..., 0], xyz[..., 1], xyz[..., 2]
    R2 = x**2 + y**2
    zp2 = (np.sqrt(z**2 + b**2) + a) ** 2
    return -G * m / np.sqrt(R2 + zp2)


def _b2(b):
    """b^2 under the root of zeta in the derivatives; for b = 0 (Kuzmin, builtin/kuzmin.py:82-84: |z| in place of
    zeta) the smallest normal number, which gives the zero in-plane z-force the reference's autodiff of |z| gives."""
    return b * b if b * b != 0.0 else TINY


def gradient_mn(G, m, a, b, xyz):
    x, y, z = xyz[..., 0], xyz[..., 1], xyz[..., 2]
    zeta = np.sqrt(z**2 + _b2(b))
    D2 = x**2 + y**2 + (a + zeta) ** 2
    f = G * m / (D2 * np.sqrt(D2))
    return np.stack([f * x, f * y, f * z * (a + zeta) / zeta], axis=-1)


def hessian_mn(G, m, a, b, xyz):
    x, y, z = xyz[..., 0], xyz[..., 1], xyz[..., 2]
    zeta = np.sqrt(z**2 + _b2(b))
    D2 = x**2 + y**2 + (a + zeta) ** 2
    D = np.sqrt(D2)
    f3 = G * m / (D2 * D)
    f5 = 3.0 * G * m / (D2 * D2 * D)
    uz = z * (a + zeta) / zeta
    duz = 1.0 + (a * b**2 / zeta**3 if b != 0.0 else 0.0)  # d(uz)/dz
    u = np.stack([x, y, uz], axis=-1)
    H = -f5[..., None, None] * (u[..., :, None] * u[..., None, :])
    H[..., 0, 0] += f3
    H[..., 1, 1] += f3
    H[..., 2, 2] += f3 * duz
    return H


# ----------------------------------------------------------------------------------------
# Hernquist (builtin/hernquist.py:79-82)


def potential_hernquist(G, m, c, xyz):
    r = _r_safe(xyz)
    return -G * m / (r + c)


def gradient_hernquist(G, m, c, xyz):
    r = _r_safe(xyz)
    return _spherical_grad(xyz, r, G * m / (r + c) ** 2)


def hessian_hernquist(G, m, c, xyz):
    r = _r_safe(xyz)
    return _spherical_hess(xyz, r, G * m / (r + c) ** 2, -2.0 * G * m / (r + c) ** 3)


# ----------------------------------------------------------------------------------------
# NFW (builtin/nfw/base.py:326-338)


def potential_nfw(G, m, r_s, xyz):
    r = _r_safe(xyz)
    x = r / r_s
    phi0 = -G * m / r_s
    return phi0 * np.log1p(x) / x


def _nfw_menc_shape(s):
    """ln(1+s) - s/(1+s), with the small-s series where the difference cancels."""
    s = np.asarray(s, dtype=np.float64)
    direct = np.log1p(s) - s / (1.0 + s)
    # series  sum_{k>=2} (-1)^k (k-1)/k s^k
    k = np.arange(2, 19)
    coeff = ((-1.0) ** k) * (k - 1.0) / k
    ser = np.zeros_like(s)
    for c in coeff[::-1]:
        ser = ser * s + c
    ser = ser * s * s
    return np.where(s < 0.0625, ser, direct)  # 0.0625^17 < 4e-21; above, ln(1+s)/m(s) < 33: no digits to lose


def gradient_nfw(G, m, r_s, xyz):
    r = _r_safe(xyz)
    s = r / r_s
    return _spherical_grad(xyz, r, G * m * _nfw_menc_shape(s) / r**2)


def hessian_nfw(G, m, r_s, xyz):
    r = _r_safe(xyz)
    s = r / r_s
    mm = _nfw_menc_shape(s)
    d1 = G * m * mm / r**2
    d2 = G * m * (s / (r_s * (1.0 + s) ** 2 * r**2) - 2.0 * mm / r**3)
    return _spherical_hess(xyz, r, d1, d2)


# ----------------------------------------------------------------------------------------
# PowerLawCutoff (builtin/powerlawcutoff.py:88-117)


def _gamma_inc_lower(a, x):
    """``gammainc(a, x) * gamma(a)`` = the unregularised lower incomplete gamma (:30-32)."""
    return sps.gammainc(a, x) * sps.gamma(a)


def potential_plc(G, m, alpha, r_c, xyz):
    r = _r_safe(xyz)
    alpha_half = alpha / 2
    s2 = (r / r_c) ** 2
    GM = G * m
    gamma_arg = 1.5 - alpha_half
    term1 = GM * _gamma_inc_lower(gamma_arg, s2) * (alpha_half - 1.5) / (r * sps.gamma(2.5 - alpha_half))
    term2 = GM * _gamma_inc_lower(1 - alpha_half, s2) / (r_c * sps.gamma(gamma_arg))
    phi_inf = GM * sps.gamma(1 - alpha_half) / (r_c * sps.gamma(gamma_arg)) if gamma_arg > 0 else 0.0
    return term1 + term2 - phi_inf


def gradient_plc(G, m, alpha, r_c, xyz):
    # d/dr of the three terms collapses to  G M P(a, s^2) / r^2,  a = 3/2 - alpha/2
    r = _r_safe(xyz)
    a = 1.5 - alpha / 2
    P = sps.gammainc(a, (r / r_c) ** 2)
    return _spherical_grad(xyz, r, G * m * P / r**2)


def hessian_plc(G, m, alpha, r_c, xyz):
    r = _r_safe(xyz)
    a = 1.5 - alpha / 2
    s2 = (r / r_c) ** 2
    P = sps.gammainc(a, s2)
    dP = s2 ** (a - 1.0) * np.exp(-s2) / sps.gamma(a)  # dP/d(s^2)
    d1 = G * m * P / r**2
    d2 = G * m * (dP * 2.0 * r / r_c**2 / r**2 - 2.0 * P / r**3)
    return _spherical_hess(xyz, r, d1, d2)


# ----------------------------------------------------------------------------------------
# (LMJ09) Logarithmic (builtin/logarithmic.py:45-108); G is unused (the depth is set by v_c)


def _log_matrix(q1, q2, q3, phi):
    sp, cp = np.sin(phi), np.cos(phi)
    R = np.array([[cp, sp, 0.0], [-sp, cp, 0.0], [0.0, 0.0, 1.0]])
    return R.T @ np.diag([1 / q1**2, 1 / q2**2, 1 / q3**2]) @ R


def potential_log(G, vc, rs, q1, q2, q3, phi, xyz):
    sp, cp = np.sin(phi), np.cos(phi)
    x = xyz[..., 0] * cp + xyz[..., 1] * sp
    y = -xyz[..., 0] * sp + xyz[..., 1] * cp
    r2 = (x / q1) ** 2 + (y / q2) ** 2 + (xyz[..., 2] / q3) ** 2
    return 0.5 * vc**2 * np.log(rs**2 + r2)


def gradient_log(G, vc, rs, q1, q2, q3, phi, xyz):
    M = _log_matrix(q1, q2, q3, phi)
    Mx = xyz @ M
    D = rs**2 + np.sum(xyz * Mx, axis=-1)
    return vc**2 * Mx / D[..., None]


def hessian_log(G, vc, rs, q1, q2, q3, phi, xyz):
    M = _log_matrix(q1, q2, q3, phi)
    Mx = xyz @ M
    D = rs**2 + np.sum(xyz * Mx, axis=-1)
    return vc**2 * (M / D[..., None, None] - 2.0 * Mx[..., :, None] * Mx[..., None, :] / (D**2)[..., None, None])


# Isochrone (builtin/isochrone.py:80-90)


def potential_iso(G, m, b, xyz):
    r = _r_safe(xyz)
    return -G * m / (b + np.sqrt(r**2 + b**2))


def gradient_iso(G, m, b, xyz):
    r = _r_safe(xyz)
    a = np.sqrt(r**2 + b**2)
    return _spherical_grad(xyz, r, G * m * r / (a * (b + a) ** 2))


def hessian_iso(G, m, b, xyz):
    r = _r_safe(xyz)
    a = np.sqrt(r**2 + b**2)
    d1 = G * m * r / (a * (b + a) ** 2)
    d2 = G * m * (1 / (a * (b + a) ** 2) - r**2 / (a**3 * (b + a) ** 2) - 2 * r**2 / (a**2 * (b + a) ** 3))
    return _spherical_hess(xyz, r, d1, d2)


# Satoh (builtin/satoh.py:63-70)


def potential_satoh(G, m, a, b, xyz):
    x, y, z = xyz[..., 0], xyz[..., 1], xyz[..., 2]
    return -G * m / np.sqrt(x**2 + y**2 + z**2 + a * (a + 2 * np.sqrt(z**2 + b**2)))


def gradient_satoh(G, m, a, b, xyz):
    x, y, z = xyz[..., 0], xyz[..., 1], xyz[..., 2]
    zeta = np.sqrt(z**2 + _b2(b))
    S = x**2 + y**2 + z**2 + a * (a + 2 * zeta)
    f = G * m / (S * np.sqrt(S))
    return np.stack([f * x, f * y, f * z * (1 + a / zeta)], axis=-1)


def hessian_satoh(G, m, a, b, xyz):
    x, y, z = xyz[..., 0], xyz[..., 1], xyz[..., 2]
    zeta = np.sqrt(z**2 + _b2(b))
    S = x**2 + y**2 + z**2 + a * (a + 2 * zeta)
    f3 = G * m / (S * np.sqrt(S))
    f5 = 3.0 * f3 / S
    u = np.stack([x, y, z * (1 + a / zeta)], axis=-1)
    duz = 1.0 + (a * b**2 / zeta**3 if b != 0.0 else 0.0)
    H = -f5[..., None, None] * (u[..., :, None] * u[..., None, :])
    H[..., 0, 0] += f3
    H[..., 1, 1] += f3
    H[..., 2, 2] += f3 * duz
    return H


# ----------------------------------------------------------------------------------------
# Profiles in an (ellipsoidal) radius m: Phi = F(m), m^2 = x^2 + (y/q1)^2 + (z/q2)^2.
#   grad = (F'/m) w,  H = (F'/m) diag(1, 1/q1^2, 1/q2^2) + (F'' - F'/m)/m^2 w w^T,  w = (x, y/q1^2, z/q2^2)


def _ellipsoidal(xyz, q1, q2):
    x, y, z = xyz[..., 0], xyz[..., 1], xyz[..., 2]
    m = np.sqrt(x**2 + (y / q1) ** 2 + (z / q2) ** 2 + TINY)
    w = np.stack([x, y / q1**2, z / q2**2], axis=-1)
    return m, w


def _ellipsoidal_grad(xyz, q1, q2, d1_of_m):
    m, w = _ellipsoidal(xyz, q1, q2)
    return (d1_of_m(m) / m)[..., None] * w


def _ellipsoidal_hess(xyz, q1, q2, d1_of_m, d2_of_m):
    m, w = _ellipsoidal(xyz, q1, q2)
    d1m = d1_of_m(m) / m
    H = ((d2_of_m(m) - d1m) / m**2)[..., None, None] * (w[..., :, None] * w[..., None, :])
    for k, s in enumerate((1.0, 1.0 / q1**2, 1.0 / q2**2)):
        H[..., k, k] += d1m * s
    return H


# TriaxialHernquist (builtin/hernquist.py:160-176): the spherical Hernquist profile at the ellipsoidal radius


def potential_thern(G, m_tot, c, q1, q2, xyz):
    m, _ = _ellipsoidal(xyz, q1, q2)
    return -G * m_tot / (m + c)


def gradient_thern(G, m_tot, c, q1, q2, xyz):
    return _ellipsoidal_grad(xyz, q1, q2, lambda m: G * m_tot / (m + c) ** 2)


def hessian_thern(G, m_tot, c, q1, q2, xyz):
    return _ellipsoidal_hess(xyz, q1, q2, lambda m: G * m_tot / (m + c) ** 2, lambda m: -2 * G * m_tot / (m + c) ** 3)


# Jaffe (builtin/jaffe.py:50-60): Phi = -GM/r_s ln(1 + r_s/r)


def potential_jaffe(G, m_tot, a, xyz):
    r = _r_safe(xyz)
    return -G * m_tot / a * np.log(1 + a / r)


def gradient_jaffe(G, m_tot, a, xyz):
    return _ellipsoidal_grad(xyz, 1.0, 1.0, lambda r: G * m_tot / (r * (r + a)))


def hessian_jaffe(G, m_tot, a, xyz):
    return _ellipsoidal_hess(xyz, 1.0, 1.0, lambda r: G * m_tot / (r * (r + a)),
                             lambda r: -G * m_tot * (2 * r + a) / (r * (r + a)) ** 2)


# Burkert (builtin/burkert.py:197-227): Phi' = G M(<r)/r^2, M(<r) = m/C [2 ln(1+x) + ln(1+x^2) - 2 atan x]
BURKERT_C = 3 * np.log(2.0) - np.pi / 2


def potential_burkert(G, m, rs, xyz):
    x = _r_safe(xyz) / rs
    xi = 1 / x
    return -G * m / (rs * BURKERT_C) * (np.pi - 2 * (1 + xi) * np.arctan(x) + 2 * (1 + xi) * np.log1p(x)
                                        - (1 - xi) * np.log1p(x**2))


def _burkert_B(s):
    """2 ln(1+s) + ln(1+s^2) - 2 atan(s); the terms cancel to O(s^3), so below s = 0.3 integrate the series of
    B' = 4 s^2 (1 - s)/(1 - s^4) instead: B = 4 s^3 sum_n s^(4n) (1/(4n+3) - s/(4n+4)), n = 0..8
    (an accuracy device shared by the C oracle and the kernels)."""
    s = np.asarray(s, dtype=np.float64)
    s4 = s**4
    e = np.zeros_like(s)
    for n in range(8, -1, -1):
        e = e * s4 + (1.0 / (4 * n + 3) - s / (4 * n + 4))
    return np.where(s < 0.3, 4.0 * s**3 * e, 2 * np.log1p(s) + np.log1p(s**2) - 2 * np.arctan(s))


def _burkert_d1(G, m, rs):
    return lambda r: G * m / BURKERT_C * _burkert_B(r / rs) / r**2


def _burkert_d2(G, m, rs):
    def f(r):
        x = r / rs
        return G * m / BURKERT_C * 4 * x**2 / ((1 + x) * (1 + x**2) * rs * r**2) - 2 * _burkert_d1(G, m, rs)(r) / r
    return f


def gradient_burkert(G, m, rs, xyz):
    return _ellipsoidal_grad(xyz, 1.0, 1.0, _burkert_d1(G, m, rs))


def hessian_burkert(G, m, rs, xyz):
    return _ellipsoidal_hess(xyz, 1.0, 1.0, _burkert_d1(G, m, rs), _burkert_d2(G, m, rs))


# StoneOstriker15 (builtin/stoneostriker15.py:150-160): Phi' = G M(<r)/r^2,
# M(<r) = 2 M/(pi (r_h - r_c)) (r_h atan(r/r_h) - r_c atan(r/r_c))


def potential_stone(G, m_tot, rc, rh, xyz):
    r = _r_safe(xyz)
    A = -2 * G * m_tot / (np.pi * (rh - rc))
    return A * ((rh * np.arctan2(r, rh) - rc * np.arctan2(r, rc)) / r + 0.5 * np.log((r**2 + rh**2) / (r**2 + rc**2)))


def _stone_T(r, rc, rh):
    """r_h atan(r/r_h) - r_c atan(r/r_c); Taylor series (16 terms) below r = 0.3 r_c where the two terms cancel
    to O(r^3).  Needs r_c < r_h (the reference's convention) for the series' convergence."""
    r = np.asarray(r, dtype=np.float64)
    uc, uh = (r / rc) ** 2, (r / rh) ** 2
    acc = np.zeros_like(r)
    pc, ph = np.ones_like(r), np.ones_like(r)
    for k in range(1, 17):
        pc, ph = pc * uc, ph * uh
        term = (ph - pc) / (2 * k + 1)
        acc = acc - term if k % 2 else acc + term
    return np.where(r < 0.3 * rc, r * acc, rh * np.arctan2(r, rh) - rc * np.arctan2(r, rc))


def _stone_d1(G, m_tot, rc, rh):
    A = 2 * G * m_tot / (np.pi * (rh - rc))
    return lambda r: A * _stone_T(r, rc, rh) / r**2


def _stone_d2(G, m_tot, rc, rh):
    A = 2 * G * m_tot / (np.pi * (rh - rc))
    # r_h^2/(r^2+r_h^2) - r_c^2/(r^2+r_c^2) = r^2 (r_h^2 - r_c^2)/((r^2+r_h^2)(r^2+r_c^2)): no cancellation
    return lambda r: A * (rh**2 - rc**2) / ((r**2 + rh**2) * (r**2 + rc**2)) - 2 * _stone_d1(G, m_tot, rc, rh)(r) / r


def gradient_stone(G, m_tot, rc, rh, xyz):
    return _ellipsoidal_grad(xyz, 1.0, 1.0, _stone_d1(G, m_tot, rc, rh))


def hessian_stone(G, m_tot, rc, rh, xyz):
    return _ellipsoidal_hess(xyz, 1.0, 1.0, _stone_d1(G, m_tot, rc, rh), _stone_d2(G, m_tot, rc, rh))


# HarmonicOscillator (builtin/example.py:75-85): Phi = 1/2 sum (omega_i x_i)^2


def potential_harmonic(G, wx, wy, wz, xyz):
    w = np.array([wx, wy, wz])
    return 0.5 * np.sum((w * xyz) ** 2, axis=-1)


def gradient_harmonic(G, wx, wy, wz, xyz):
    return np.array([wx, wy, wz]) ** 2 * xyz


def hessian_harmonic(G, wx, wy, wz, xyz):
    return np.broadcast_to(np.diag(np.array([wx, wy, wz]) ** 2), xyz.shape[:-1] + (3, 3)).copy()


# HenonHeiles (builtin/example.py:159-176): Phi = (R^2/2 + coeff (x^2 y - y^3/3)) / timescale^2


def potential_henon(G, coeff, ts, xyz):
    x, y = xyz[..., 0], xyz[..., 1]
    return ((x**2 + y**2) / 2 + coeff * (x**2 * y - y**3 / 3.0)) / ts**2


def gradient_henon(G, coeff, ts, xyz):
    x, y = xyz[..., 0], xyz[..., 1]
    return np.stack([x + 2 * coeff * x * y, y + coeff * (x**2 - y**2), np.zeros_like(x)], axis=-1) / ts**2


def hessian_henon(G, coeff, ts, xyz):
    x, y = xyz[..., 0], xyz[..., 1]
    H = np.zeros(xyz.shape[:-1] + (3, 3))
    H[..., 0, 0] = 1 + 2 * coeff * y
    H[..., 0, 1] = H[..., 1, 0] = 2 * coeff * x
    H[..., 1, 1] = 1 - 2 * coeff * y
    return H / ts**2


_NEW = {KIND_TRIAXIAL_HERNQUIST: (potential_thern, gradient_thern, hessian_thern),
        KIND_JAFFE: (potential_jaffe, gradient_jaffe, hessian_jaffe),
        KIND_BURKERT: (potential_burkert, gradient_burkert, hessian_burkert),
        KIND_STONE: (potential_stone, gradient_stone, hessian_stone),
        KIND_HARMONIC: (potential_harmonic, gradient_harmonic, hessian_harmonic),
        KIND_HENON_HEILES: (potential_henon, gradient_henon, hessian_henon)}

_POT = {KIND_LOG: potential_log, KIND_ISOCHRONE: potential_iso, KIND_SATOH: potential_satoh, KIND_MN: potential_mn, KIND_HERNQUIST: potential_hernquist, KIND_NFW: potential_nfw, KIND_PLC: potential_plc}
_GRAD = {KIND_LOG: gradient_log, KIND_ISOCHRONE: gradient_iso, KIND_SATOH: gradient_satoh, KIND_MN: gradient_mn, KIND_HERNQUIST: gradient_hernquist, KIND_NFW: gradient_nfw, KIND_PLC: gradient_plc}
_HESS = {KIND_LOG: hessian_log, KIND_ISOCHRONE: hessian_iso, KIND_SATOH: hessian_satoh, KIND_MN: hessian_mn, KIND_HERNQUIST: hessian_hernquist, KIND_NFW: hessian_nfw, KIND_PLC: hessian_plc}


for _k, (_p, _g, _h) in _NEW.items():
    _POT[_k], _GRAD[_k], _HESS[_k] = _p, _g, _h

# ----------------------------------------------------------------------------------------
# composite evaluation, summed in component order (base_multi.py:39-82)


def _sum_components(table, pot: Potential, xyz, t=0.0):
    xyz = np.asarray(xyz, dtype=np.float64)
    total = None
    for grp in pot.group_list():
        sub = None
        for i in grp:
            c = pot.components[i]
            v = table[c.kind](pot.G, *c.params_at(float(t)), xyz)
            sub = v if sub is None else sub + v
        total = sub if total is None else total + sub
    return total


def potential(pot: Potential, xyz, t=0.0):
    return _sum_components(_POT, pot, xyz, t)


def gradient(pot: Potential, xyz, t=0.0):
    return _sum_components(_GRAD, pot, xyz, t)


def acceleration(pot: Potential, xyz, t=0.0):
    """``acceleration = -gradient`` (register_funcs.py:327-340)."""
    return -gradient(pot, xyz, t)


def hessian(pot: Potential, xyz, t=0.0):
    return _sum_components(_HESS, pot, xyz, t)


def laplacian(pot: Potential, xyz, t=0.0):
    """trace of the Hessian (base.py:191-200)."""
    return np.trace(hessian(pot, xyz, t), axis1=-2, axis2=-1)


def density(pot: Potential, xyz, t=0.0):
    """laplacian / (4 pi G) (base.py:212-218)."""
    return laplacian(pot, xyz, t) / (4 * np.pi * pot.G)


def tidal_tensor(pot: Potential, xyz, t=0.0):
    """H - tr(H)/3 I (register_funcs.py:347-377)."""
    H = hessian(pot, xyz, t)
    tr = np.trace(H, axis1=-2, axis2=-1)
    return H - np.eye(3) * (tr / 3)[..., None, None]


def d2potential_dr2(pot: Potential, xyz, t=0.0):
    """rhat . H . rhat (register_funcs.py:442-457); rhat = xyz/|xyz| (plain norm)."""
    xyz = np.asarray(xyz, dtype=np.float64)
    rhat = xyz / np.linalg.norm(xyz, axis=-1, keepdims=True)
    H = hessian(pot, xyz, t)
    return np.einsum("...i,...ij,...j->...", rhat, H, rhat)


def dpotential_dr(pot: Potential, xyz, t=0.0):
    """rhat . grad (register_funcs.py:415-426)."""
    xyz = np.asarray(xyz, dtype=np.float64)
    rhat = xyz / np.linalg.norm(xyz, axis=-1, keepdims=True)
    return np.sum(gradient(pot, xyz, t) * rhat, axis=-1)


# ----------------------------------------------------------------------------------------
# MN3 host-side parameter fit (builtin/mn3.py:31-52,90-119)

MN3_K_POS_DENS = np.array(
    [
        [0.0036, -0.0330, 0.1117, -0.1335, 0.1749],
        [-0.0131, 0.1090, -0.3035, 0.2921, -5.7976],
        [-0.0048, 0.0454, -0.1425, 0.1012, 6.7120],
        [-0.0158, 0.0993, -0.2070, -0.7089, 0.6445],
        [-0.0319, 0.1514, -0.1279, -0.9325, 2.6836],
        [-0.0326, 0.1816, -0.2943, -0.6329, 2.3193],
    ]
)
MN3_K_NEG_DENS = np.array(
    [
        [-0.0090, 0.0640, -0.1653, 0.1164, 1.9487],
        [0.0173, -0.0903, 0.0877, 0.2029, -1.3077],
        [-0.0051, 0.0287, -0.0361, -0.0544, 0.2242],
        [-0.0358, 0.2610, -0.6987, -0.1193, 2.0074],
        [-0.0830, 0.4992, -0.7967, -1.2966, 4.4441],
        [-0.0247, 0.1718, -0.4124, -0.5944, 0.7333],
    ]
)
MN3_B_COEFFS_EXP = np.array([-0.269, 1.08, 1.092])
MN3_B_COEFFS_SECH2 = np.array([-0.033, 0.262, 0.659])


def mn3_components(m_tot, h_R, h_z, *, sech2: bool, positive_density: bool) -> list[Component]:
    hzR = h_z / h_R
    K = MN3_K_POS_DENS if positive_density else MN3_K_NEG_DENS
    bco = MN3_B_COEFFS_SECH2 if sech2 else MN3_B_COEFFS_EXP
    b_hR = bco @ np.array([hzR**3, hzR**2, hzR])
    x = np.vander(np.array([b_hR]), N=5)[0]
    pv = K @ x
    ms = pv[:3] * m_tot
    as_ = pv[3:] * h_R
    b = b_hR * h_R
    return [Component(KIND_MN, (float(ms[i]), float(as_[i]), float(b)), f"disk.mn{i}") for i in range(3)]


# ----------------------------------------------------------------------------------------
# the three Milky-Way models (builtin/milkyway.py:65-97,203-236,275-313)


def milky_way_potential(G=G_GALACTIC) -> Potential:
    comps = (
        Component(KIND_MN, (6.8e10, 3.0, 0.28), "disk"),
        Component(KIND_NFW, (5.4e11, 15.62), "halo"),
        Component(KIND_HERNQUIST, (5e9, 1.0), "bulge"),
        Component(KIND_HERNQUIST, (1.71e9, 0.07), "nucleus"),
    )
    return Potential(comps, G, name="MilkyWayPotential")


def milky_way_potential_2022(G=G_GALACTIC) -> Potential:
    disk = mn3_components(4.7717e10, 2.6, 0.3, sech2=True, positive_density=True)
    comps = (
        *disk,
        Component(KIND_NFW, (5.5427e11, 15.626), "halo"),
        Component(KIND_HERNQUIST, (5e9, 1.0), "bulge"),
        Component(KIND_HERNQUIST, (1.8142e9, 68.8867 * 0.001), "nucleus"),
    )
    return Potential(comps, G, groups=((0, 1, 2), (3,), (4,), (5,)), name="MilkyWayPotential2022")


def bovy_mw_potential_2014(G=G_GALACTIC) -> Potential:
    comps = (
        Component(KIND_MN, (68_193_902_782.346756, 3.0, 280 * 0.001), "disk"),
        Component(KIND_PLC, (4501365375.06545, 1.8, 1.9), "bulge"),
        Component(KIND_NFW, (4.3683325e11, 16.0), "halo"),
    )
    return Potential(comps, G, name="BovyMWPotential2014")


KMS = 0.001022712165045695  # 1 km/s in kpc/Myr


def lm10_potential(G=G_GALACTIC) -> Potential:
    """builtin/milkyway.py:101-169 (Law & Majewski 2010)."""
    comps = (
        Component(KIND_MN, (1e11, 6.5, 0.26), "disk"),
        Component(KIND_HERNQUIST, (3.4e10, 0.7), "bulge"),
        Component(KIND_LOG, (np.sqrt(2.0) * 121.858 * KMS, 12.0, 1.38, 1.0, 1.36, np.deg2rad(97.0)), "halo"),
    )
    return Potential(comps, G, name="LM10Potential")


def single(kind: int, *params: float, G=G_GALACTIC) -> Potential:
    return Potential((Component(kind, tuple(float(p) for p in params)),), G, name=KIND_NAMES[kind])


def mn3_potential(m_tot, h_R, h_z, *, sech2, positive_density=False, G=G_GALACTIC) -> Potential:
    comps = tuple(mn3_components(m_tot, h_R, h_z, sech2=sech2, positive_density=positive_density))
    return Potential(comps, G, groups=((0, 1, 2),), name="MN3Sech2" if sech2 else "MN3Exponential")


MODELS = {
    "MilkyWayPotential": milky_way_potential,
    "MilkyWayPotential2022": milky_way_potential_2022,
    "BovyMWPotential2014": bovy_mw_potential_2014,
}


def circular_velocity(pot: Potential, r: np.ndarray) -> np.ndarray:
    """v_c(r) = sqrt(r dPhi/dr) in the z=0 plane; used only to draw synthetic ICs."""
    xyz = np.stack([r, np.zeros_like(r), np.zeros_like(r)], axis=-1)
    g = gradient(pot, xyz)
    return np.sqrt(r * g[..., 0])
