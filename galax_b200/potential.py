"""Host-side mirror of ``galax.potential`` for the supported Milky-Way composites.

Same class names, parameter names, method names and argument meaning as the reference
(``/root/reference/src/galax/potential/_src/builtin/*.py`` and the public functions of
``potential/_src/api.py`` / ``register_funcs.py``), but every evaluation is one launch of the CUDA
library through the C ABI -- there is no CPU path.  Values are plain fp64 arrays in the potential's
unit system (galactic: kpc, Myr, Msun), i.e. what the reference returns for bare-array inputs
(``register_funcs.py:86-98``).  numpy in -> numpy out, torch in -> torch out (a CUDA tensor stays on
the device and nothing is copied).

Out of scope and rejected loudly (``NotImplementedError``): time-dependent parameters
(``potential/_src/params/core.py``), component classes other than the four below, unit systems other
than the one the parameters are given in.
"""

from __future__ import annotations

import ctypes as C
import dataclasses
import math
from typing import Any, Iterable, Sequence

import numpy as np

from . import _lib

# astropy CODATA2018 G in kpc^3 / (Msun Myr^2) -- what ``pot.constants["G"].value`` holds for
# ``units="galactic"`` (potential/_src/base.py:35,91-95).  galax objects passed through
# ``galax_b200.interop`` carry their own value instead.
G_GALACTIC = 4.498502151469553e-12
KMS = 0.001022712165045695  # 1 km/s in kpc/Myr


@dataclasses.dataclass(frozen=True)
class LinearParameter:
    """``gp.params.LinearParameter`` (potential/_src/params/core.py:25-110): p(t) = slope (t - point_time) + point_value,
    all three in the potential's units (galactic: Msun, kpc, Myr)."""

    slope: float
    point_time: float
    point_value: float

    def __call__(self, t: float) -> float:
        return self.slope * (t - self.point_time) + self.point_value


@dataclasses.dataclass(frozen=True)
class ConstantParameter:
    """``gp.params.ConstantParameter`` (potential/_src/params/constant.py)."""

    value: float

    def __call__(self, t: float = 0.0) -> float:
        return self.value


class _Param(float):
    """A parameter value at t = 0 that remembers its time derivative (0 for constants)."""

    rate: float = 0.0

    def __new__(cls, value: float, rate: float = 0.0):
        obj = super().__new__(cls, value)
        obj.rate = float(rate)
        return obj


def _const(name: str, v: Any) -> float:
    """ConstantParameter or LinearParameter; any other callable (``UserParameter``) is rejected."""
    if isinstance(v, LinearParameter):
        return _Param(v.point_value - v.slope * v.point_time, v.slope)
    if isinstance(v, ConstantParameter):
        return _Param(v.value)
    if callable(v):
        raise NotImplementedError(
            f"parameter {name!r} is a general function of time; galax_b200 kernels support ConstantParameter and "
            "LinearParameter only"
        )
    if hasattr(v, "value"):  # unxt.Quantity-like: caller must already be in the potential's units
        v = v.value
    return _Param(float(v), getattr(v, "rate", 0.0))


def _static(name: str, v: Any) -> float:
    """For parameters that enter host-side preprocessing (MN3 fits, rotation matrices, gamma functions)."""
    out = _const(name, v)
    if getattr(out, "rate", 0.0) != 0.0:
        raise NotImplementedError(f"parameter {name!r} of this potential cannot be time-dependent")
    return float(out)


class AbstractPotential:
    """Common evaluation API (reference: ``AbstractPotential`` potential/_src/base.py:43-553)."""

    G: float = G_GALACTIC
    units = "galactic"

    # -- to be provided by subclasses: list of (kind, (p0, p1, p2)) in the reference's summation order
    def _flat_components(self) -> list[tuple[int, tuple[float, ...]]]:
        raise NotImplementedError

    # ---------------------------------------------------------------- C ABI marshalling
    def c_struct(self) -> _lib.GxPotential:
        comps = self._flat_components()
        if len(comps) > _lib.GX_MAX_COMPONENTS:
            raise NotImplementedError(f"at most {_lib.GX_MAX_COMPONENTS} components are supported")
        P = _lib.GxPotential()
        P.n = len(comps)
        P.G = float(self.G)
        groups = self._flat_groups()
        for i, (kind, params) in enumerate(comps):
            P.c[i].kind = kind
            P.c[i].reserved = groups[i]
            for j, v in enumerate(params):
                P.c[i].p[j] = float(v)
                P.c[i].dp[j] = float(getattr(v, "rate", 0.0))
        return P

    def _flat_groups(self) -> list[int]:
        """Summation group of every flat component (``gx_component.reserved``): components that were ONE reference
        component (the three Miyamoto-Nagai terms of an MN3 disk, a nested composite) share a non-zero id and the
        reference-order kernels (``GX_SCHEME_STRICT``) sum them first, as base_multi.py:48-55 does.  0 = its own."""
        return [0] * len(self._flat_components())

    @property
    def is_time_dependent(self) -> bool:
        return any(getattr(v, "rate", 0.0) != 0.0 for _, params in self._flat_components() for v in params)

    # ---------------------------------------------------------------- evaluation
    def _eval(self, q, t, what: int):
        torch = _lib.require_cuda()
        dev_q, restore = _to_device(q)
        batch = dev_q.shape[:-1]
        if dev_q.shape[-1] != 3:
            raise ValueError("positions must have a trailing axis of length 3")
        xyz = dev_q.reshape(-1, 3).contiguous()
        N = xyz.shape[0]
        out = {}
        ptr = {1: None, 2: None, 4: None, 8: None}
        if what & _lib.PHI:
            out["phi"] = torch.empty((N,), dtype=torch.float64, device=xyz.device)
            ptr[1] = out["phi"].data_ptr()
        if what & _lib.GRAD:
            out["grad"] = torch.empty((N, 3), dtype=torch.float64, device=xyz.device)
            ptr[2] = out["grad"].data_ptr()
        if what & _lib.ACC:
            out["acc"] = torch.empty((N, 3), dtype=torch.float64, device=xyz.device)
            ptr[4] = out["acc"].data_ptr()
        if what & _lib.HESS:
            out["hess"] = torch.empty((N, 3, 3), dtype=torch.float64, device=xyz.device)
            ptr[8] = out["hess"].data_ptr()
        P = self.c_struct()
        with torch.cuda.device(xyz.device):
            stream = torch.cuda.current_stream().cuda_stream
            rc = _lib.lib().gx_potential_eval(
                C.byref(P), xyz.data_ptr(), float(_scalar_time(t)), N, what, ptr[1], ptr[2], ptr[4], ptr[8], stream
            )
        _lib.check(rc, "gx_potential_eval")
        shapes = {"phi": (), "grad": (3,), "acc": (3,), "hess": (3, 3)}
        return {k: restore(v.reshape(tuple(batch) + shapes[k])) for k, v in out.items()}

    def potential(self, q, t=0.0):
        """``pot.potential(q, t)`` (api.py:27-110, register_funcs.py:33-80)."""
        return self._eval(q, t, _lib.PHI)["phi"]

    def gradient(self, q, t=0.0):
        """``pot.gradient(q, t)`` (register_funcs.py:86-155)."""
        return self._eval(q, t, _lib.GRAD)["grad"]

    def acceleration(self, q, t=0.0):
        """``pot.acceleration(q, t)`` = -gradient (register_funcs.py:327-340)."""
        return self._eval(q, t, _lib.ACC)["acc"]

    def hessian(self, q, t=0.0):
        """``pot.hessian(q, t)`` (register_funcs.py:276-320)."""
        return self._eval(q, t, _lib.HESS)["hess"]

    def laplacian(self, q, t=0.0):
        """trace of the Hessian (base.py:191-200)."""
        H = self.hessian(q, t)
        return H[..., 0, 0] + H[..., 1, 1] + H[..., 2, 2]

    def density(self, q, t=0.0):
        """laplacian / (4 pi G) (base.py:212-218)."""
        return self.laplacian(q, t) / (4 * math.pi * self.G)

    def tidal_tensor(self, q, t=0.0):
        """H - tr(H)/3 I (register_funcs.py:347-377)."""
        H = self.hessian(q, t)
        tr3 = (H[..., 0, 0] + H[..., 1, 1] + H[..., 2, 2]) / 3
        eye = np.eye(3) if isinstance(H, np.ndarray) else _eye_like(H)
        return H - eye * tr3[..., None, None]

    def d2potential_dr2(self, q, t=0.0):
        """rhat . H . rhat (register_funcs.py:442-457)."""
        H = self.hessian(q, t)
        qq = q if not isinstance(H, np.ndarray) else np.asarray(q, dtype=np.float64)
        rhat = qq / ((qq * qq).sum(-1, keepdims=True) ** 0.5)
        return ((H * rhat[..., None, :]).sum(-1) * rhat).sum(-1)

    def dpotential_dr(self, q, t=0.0):
        """rhat . grad (register_funcs.py:415-426)."""
        g = self.gradient(q, t)
        qq = q if not isinstance(g, np.ndarray) else np.asarray(q, dtype=np.float64)
        rhat = qq / ((qq * qq).sum(-1, keepdims=True) ** 0.5)
        return (g * rhat).sum(-1)

    def local_circular_velocity(self, q, t=0.0):
        """sqrt(r |dPhi/dr|) (register_funcs.py:384-400)."""
        qq = q if not isinstance(q, (list, tuple)) else np.asarray(q, dtype=np.float64)
        r = (qq * qq).sum(-1) ** 0.5
        return (r * abs(self.dpotential_dr(qq, t))) ** 0.5

    def spherical_mass_enclosed(self, q, t=0.0):
        """r^2 |dPhi/dr| / G (register_funcs.py:470-490)."""
        qq = q if not isinstance(q, (list, tuple)) else np.asarray(q, dtype=np.float64)
        return (qq * qq).sum(-1) * abs(self.dpotential_dr(qq, t)) / self.G

    # ---------------------------------------------------------------- orbits (base.py:384-456)
    def evaluate_orbit(self, w0, t, **kw):
        from .dynamics import evaluate_orbit

        return evaluate_orbit(self, w0, t, **kw)

    def compute_orbit(self, w0, t, **kw):
        from .dynamics import compute_orbit

        return compute_orbit(self, w0, t, **kw)

    def __add__(self, other: "AbstractPotential") -> "CompositePotential":
        return CompositePotential(a=self, b=other)


def _scalar_time(t) -> float:
    if hasattr(t, "value"):
        t = t.value
    a = np.asarray(t, dtype=np.float64)
    return float(a.reshape(-1)[0]) if a.size else 0.0


def _eye_like(H):
    import torch

    return torch.eye(3, dtype=H.dtype, device=H.device)


def _to_host(r):
    """Device result -> host tensor through PINNED memory.

    Measured on the B200 box: a 58 MB device-to-host copy takes 14.5 ms into pageable memory but 1.05 ms into
    pinned memory; torch's caching host allocator recycles pinned blocks, so after the first call of a given size
    the allocation is free."""
    import torch

    out = torch.empty(r.shape, dtype=r.dtype, pin_memory=True)
    out.copy_(r, non_blocking=True)
    torch.cuda.current_stream(r.device).synchronize()
    return out


def _to_host_many(tensors):
    """Several device results -> pinned host tensors: all copies queued, one synchronisation."""
    import torch

    outs = [torch.empty(r.shape, dtype=r.dtype, pin_memory=True) for r in tensors]
    for o, r in zip(outs, tensors):
        o.copy_(r, non_blocking=True)
    if tensors:
        torch.cuda.current_stream(tensors[0].device).synchronize()
    return outs


def _to_device(x):
    """-> (float64 CUDA tensor, function mapping a CUDA result back to the caller's array kind)."""
    import torch

    if isinstance(x, torch.Tensor):
        if x.is_cuda:
            return x.to(torch.float64), (lambda r: r)
        return x.to(device="cuda", dtype=torch.float64, non_blocking=x.is_pinned()), _to_host
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    d = torch.from_numpy(a).to("cuda", non_blocking=False)
    return d, (lambda r: _to_host(r).numpy())


# -----------------------------------------------------------------------------------------------
# single components


@dataclasses.dataclass(frozen=True)
class MiyamotoNagaiPotential(AbstractPotential):
    """builtin/miyamotonagai.py: Phi = -G m / sqrt(R^2 + (a + sqrt(z^2 + b^2))^2)."""

    m_tot: float
    a: float
    b: float
    G: float = G_GALACTIC

    def _flat_components(self):
        return [(_lib.KIND_MN, (_const("m_tot", self.m_tot), _const("a", self.a), _const("b", self.b)))]


@dataclasses.dataclass(frozen=True)
class HernquistPotential(AbstractPotential):
    """builtin/hernquist.py: Phi = -G m / (r + r_s)."""

    m_tot: float
    r_s: float
    G: float = G_GALACTIC

    def _flat_components(self):
        return [(_lib.KIND_HERNQUIST, (_const("m_tot", self.m_tot), _const("r_s", self.r_s)))]


@dataclasses.dataclass(frozen=True)
class KeplerPotential(AbstractPotential):
    """builtin/kepler.py: Phi = -G m / r  (evaluated as a Hernquist sphere with r_s = 0)."""

    m_tot: float
    G: float = G_GALACTIC

    def _flat_components(self):
        return [(_lib.KIND_HERNQUIST, (_const("m_tot", self.m_tot), 0.0))]


@dataclasses.dataclass(frozen=True)
class PlummerPotential(AbstractPotential):
    """builtin/plummer.py: Phi = -G m / sqrt(r^2 + r_s^2)  (a Miyamoto-Nagai model with a = 0)."""

    m_tot: float
    r_s: float
    G: float = G_GALACTIC

    def _flat_components(self):
        return [(_lib.KIND_MN, (_const("m_tot", self.m_tot), 0.0, _const("r_s", self.r_s)))]


@dataclasses.dataclass(frozen=True)
class KuzminPotential(AbstractPotential):
    """builtin/kuzmin.py: Phi = -G m / sqrt(R^2 + (r_s + |z|)^2)  (a Miyamoto-Nagai model with b = 0)."""

    m_tot: float
    r_s: float
    G: float = G_GALACTIC

    def _flat_components(self):
        return [(_lib.KIND_MN, (_const("m_tot", self.m_tot), _const("r_s", self.r_s), 0.0))]


@dataclasses.dataclass(frozen=True)
class IsochronePotential(AbstractPotential):
    """builtin/isochrone.py: Phi = -G m / (r_s + sqrt(r^2 + r_s^2))."""

    m_tot: float
    r_s: float
    G: float = G_GALACTIC

    def _flat_components(self):
        return [(_lib.KIND_ISOCHRONE, (_const("m_tot", self.m_tot), _const("r_s", self.r_s)))]


@dataclasses.dataclass(frozen=True)
class SatohPotential(AbstractPotential):
    """builtin/satoh.py: Phi = -G m / sqrt(R^2 + z^2 + a (a + 2 sqrt(z^2 + b^2)))."""

    m_tot: float
    a: float
    b: float
    G: float = G_GALACTIC

    def _flat_components(self):
        return [(_lib.KIND_SATOH, (_const("m_tot", self.m_tot), _const("a", self.a), _const("b", self.b)))]


@dataclasses.dataclass(frozen=True)
class TriaxialHernquistPotential(AbstractPotential):
    """builtin/hernquist.py:89-176: the Hernquist profile at m^2 = x^2 + (y/q1)^2 + (z/q2)^2."""

    m_tot: float
    r_s: float
    q1: float = 1.0
    q2: float = 1.0
    G: float = G_GALACTIC

    def _flat_components(self):
        return [(_lib.KIND_TRIAXIAL_HERNQUIST, tuple(_const(n, getattr(self, n)) for n in ("m_tot", "r_s", "q1", "q2")))]


@dataclasses.dataclass(frozen=True)
class JaffePotential(AbstractPotential):
    """builtin/jaffe.py: Phi = -G m / r_s ln(1 + r_s / r)."""

    m_tot: float
    r_s: float
    G: float = G_GALACTIC

    def _flat_components(self):
        return [(_lib.KIND_JAFFE, (_const("m_tot", self.m_tot), _const("r_s", self.r_s)))]


@dataclasses.dataclass(frozen=True)
class BurkertPotential(AbstractPotential):
    """builtin/burkert.py:37-227: cored halo; ``m`` is the characteristic mass pi rho_0 r_s^3 (3 ln 2 - pi/2)."""

    m: float
    r_s: float
    G: float = G_GALACTIC

    def _flat_components(self):
        return [(_lib.KIND_BURKERT, (_const("m", self.m), _const("r_s", self.r_s)))]

    @classmethod
    def from_central_density(cls, rho_0: float, r_s: float, **kw) -> "BurkertPotential":
        """builtin/burkert.py:114-150."""
        return cls(m=math.pi * rho_0 * r_s**3 * (3 * math.log(2.0) - math.pi / 2), r_s=r_s, **kw)

    def rho0(self) -> float:
        """Central density (builtin/burkert.py:103-109)."""
        return self.m / ((3 * math.log(2.0) - math.pi / 2) * math.pi * self.r_s**3)


@dataclasses.dataclass(frozen=True)
class StoneOstriker15Potential(AbstractPotential):
    """builtin/stoneostriker15.py:19-160: rho ~ 1 / ((1 + r^2/r_c^2)(1 + r^2/r_h^2))."""

    m_tot: float
    r_c: float
    r_h: float
    G: float = G_GALACTIC

    def _flat_components(self):
        return [(_lib.KIND_STONE, (_const("m_tot", self.m_tot), _const("r_c", self.r_c), _const("r_h", self.r_h)))]


@dataclasses.dataclass(frozen=True)
class HarmonicOscillatorPotential(AbstractPotential):
    """builtin/example.py:23-100: Phi = 1/2 sum (omega_i x_i)^2; ``omega`` a scalar or three values [1/Myr]."""

    omega: Any
    G: float = G_GALACTIC

    def _flat_components(self):
        w = np.broadcast_to(np.asarray(self.omega, dtype=np.float64), (3,))
        return [(_lib.KIND_HARMONIC, tuple(float(v) for v in w))]

    def density(self, q, t=0.0):
        """The reference's own ``_density`` (example.py:86-98): sum over ``atleast_1d(omega)``^2 / (4 pi G) -- for a
        scalar omega that is a third of laplacian / (4 pi G); mirrored as is."""
        rho = float(np.sum(np.atleast_1d(np.asarray(self.omega, dtype=np.float64)) ** 2)) / (4 * math.pi * self.G)
        lap = self.laplacian(q, t)
        return lap * 0 + rho


@dataclasses.dataclass(frozen=True)
class HenonHeilesPotential(AbstractPotential):
    """builtin/example.py:107-176: Phi = (R^2/2 + coeff (x^2 y - y^3/3)) / timescale^2."""

    coeff: float
    timescale: float
    G: float = G_GALACTIC

    def _flat_components(self):
        return [(_lib.KIND_HENON_HEILES, (_const("coeff", self.coeff), _const("timescale", self.timescale)))]


@dataclasses.dataclass(frozen=True)
class NullPotential(AbstractPotential):
    """builtin/null.py: Phi = 0 (no components reach the device; every output is zero)."""

    G: float = G_GALACTIC

    def _flat_components(self):
        return []


@dataclasses.dataclass(frozen=True)
class LMJ09LogarithmicPotential(AbstractPotential):
    """builtin/logarithmic.py:57-108: Phi = v_c^2/2 ln(r_s^2 + (x'/q1)^2 + (y'/q2)^2 + (z/q3)^2), (x', y') rotated by
    ``phi`` about z.  ``v_c`` in the potential's speed unit (kpc/Myr: multiply km/s by ``KMS``), ``phi`` in radians."""

    v_c: float
    r_s: float
    q1: float = 1.0
    q2: float = 1.0
    q3: float = 1.0
    phi: float = 0.0
    G: float = G_GALACTIC

    def _flat_components(self):
        return [(_lib.KIND_LOG, tuple(_const(n, getattr(self, n)) for n in ("v_c", "r_s", "q1", "q2", "q3", "phi")))]


@dataclasses.dataclass(frozen=True)
class LogarithmicPotential(AbstractPotential):
    """builtin/logarithmic.py:27-53: the spherical case, Phi = v_c^2/2 ln(r_s^2 + r^2)."""

    v_c: float
    r_s: float
    G: float = G_GALACTIC

    def _flat_components(self):
        return [(_lib.KIND_LOG, (_const("v_c", self.v_c), _const("r_s", self.r_s), 1.0, 1.0, 1.0, 0.0))]


@dataclasses.dataclass(frozen=True)
class NFWPotential(AbstractPotential):
    """builtin/nfw/base.py: Phi = -(G m / r_s) log(1 + r/r_s) / (r/r_s)."""

    m: float
    r_s: float
    G: float = G_GALACTIC

    def _flat_components(self):
        return [(_lib.KIND_NFW, (_const("m", self.m), _const("r_s", self.r_s)))]


@dataclasses.dataclass(frozen=True)
class PowerLawCutoffPotential(AbstractPotential):
    """builtin/powerlawcutoff.py: rho ~ r^-alpha exp(-(r/r_c)^2); 0 <= alpha < 2 supported."""

    m_tot: float
    alpha: float
    r_c: float
    G: float = G_GALACTIC

    def _flat_components(self):
        return [(_lib.KIND_PLC, (_const("m_tot", self.m_tot), _const("alpha", self.alpha), _const("r_c", self.r_c)))]


# MN3 fit tables of Smith et al. (2015) as tabulated in builtin/mn3.py:31-52
_MN3_K_POS = np.array(
    [
        [0.0036, -0.0330, 0.1117, -0.1335, 0.1749],
        [-0.0131, 0.1090, -0.3035, 0.2921, -5.7976],
        [-0.0048, 0.0454, -0.1425, 0.1012, 6.7120],
        [-0.0158, 0.0993, -0.2070, -0.7089, 0.6445],
        [-0.0319, 0.1514, -0.1279, -0.9325, 2.6836],
        [-0.0326, 0.1816, -0.2943, -0.6329, 2.3193],
    ]
)
_MN3_K_NEG = np.array(
    [
        [-0.0090, 0.0640, -0.1653, 0.1164, 1.9487],
        [0.0173, -0.0903, 0.0877, 0.2029, -1.3077],
        [-0.0051, 0.0287, -0.0361, -0.0544, 0.2242],
        [-0.0358, 0.2610, -0.6987, -0.1193, 2.0074],
        [-0.0830, 0.4992, -0.7967, -1.2966, 4.4441],
        [-0.0247, 0.1718, -0.4124, -0.5944, 0.7333],
    ]
)


@dataclasses.dataclass(frozen=True)
class _AbstractMN3Potential(AbstractPotential):
    """builtin/mn3.py:55-130: three Miyamoto-Nagai disks fitted to an exponential / sech^2 disk."""

    m_tot: float
    h_R: float
    h_z: float
    positive_density: bool = False
    G: float = G_GALACTIC

    _b_coeffs = (0.0, 0.0, 0.0)

    def _get_mn_components(self) -> list[MiyamotoNagaiPotential]:
        # builtin/mn3.py:90-119 (host-side algebra, same operation order)
        m_tot, hR, hz = _static("m_tot", self.m_tot), _static("h_R", self.h_R), _static("h_z", self.h_z)
        hzR = hz / hR
        K = _MN3_K_POS if self.positive_density else _MN3_K_NEG
        b_hR = np.asarray(self._b_coeffs) @ np.array([hzR**3, hzR**2, hzR])
        x = np.vander(np.array([b_hR]), N=5)[0]
        pv = K @ x
        ms, as_, b = pv[:3] * m_tot, pv[3:] * hR, b_hR * hR
        return [MiyamotoNagaiPotential(float(ms[i]), float(as_[i]), float(b), G=self.G) for i in range(3)]

    def _flat_components(self):
        return [c._flat_components()[0] for c in self._get_mn_components()]


@dataclasses.dataclass(frozen=True)
class MN3ExponentialPotential(_AbstractMN3Potential):
    _b_coeffs = (-0.269, 1.08, 1.092)


@dataclasses.dataclass(frozen=True)
class MN3Sech2Potential(_AbstractMN3Potential):
    _b_coeffs = (-0.033, 0.262, 0.659)


# -----------------------------------------------------------------------------------------------
# composites


class CompositePotential(AbstractPotential):
    """composite.py:32-111 / base_multi.py: named components, summed in insertion order."""

    def __init__(self, potentials: dict[str, AbstractPotential] | Iterable = (), /, G: float | None = None, **kw):
        comps = dict(potentials)
        comps.update(kw)
        for k, v in comps.items():
            if not isinstance(v, AbstractPotential):
                raise TypeError(f"component {k!r} is not a galax_b200 potential")
        self._data = comps
        gs = {float(v.G) for v in comps.values()}
        if G is None:
            if len(gs) > 1:
                raise ValueError("components disagree on G")
            G = gs.pop() if gs else G_GALACTIC
        self.G = float(G)

    def keys(self):
        return self._data.keys()

    def values(self):
        return self._data.values()

    def items(self):
        return self._data.items()

    def __getitem__(self, k):
        return self._data[k]

    def __len__(self):
        return len(self._data)

    def _flat_components(self):
        out = []
        for v in self._data.values():
            out.extend(v._flat_components())
        return out

    def _flat_groups(self):
        out: list[int] = []
        for gid, v in enumerate(self._data.values(), start=1):
            n = len(v._flat_components())
            if n > 1 and isinstance(v, CompositePotential) and len(v) > 1 and any(
                len(c._flat_components()) > 1 for c in v.values()
            ):
                out.extend([-1] * n)  # groups inside a group: one level is all gx_component.reserved can express
            else:
                out.extend([gid if n > 1 else 0] * n)
        return out


class MilkyWayPotential(CompositePotential):
    """builtin/milkyway.py:173-236 (Price-Whelan 2017 / gala ``MilkyWayPotential``)."""

    _defaults = {
        "disk": dict(m_tot=6.8e10, a=3.0, b=0.28),
        "halo": dict(m=5.4e11, r_s=15.62),
        "bulge": dict(m_tot=5e9, r_s=1.0),
        "nucleus": dict(m_tot=1.71e9, r_s=0.07),
    }
    _classes = {"disk": MiyamotoNagaiPotential, "halo": NFWPotential, "bulge": HernquistPotential,
                "nucleus": HernquistPotential}  # fmt: skip

    def __init__(self, *, G: float = G_GALACTIC, **overrides):
        comps = {}
        for name, cls in self._classes.items():
            o = overrides.pop(name, None)
            if isinstance(o, AbstractPotential):
                comps[name] = o
            else:
                params = dict(self._defaults[name])
                params.update(o or {})
                comps[name] = cls(**params, G=G)
        if overrides:
            raise TypeError(f"unknown components {sorted(overrides)}")
        super().__init__(comps, G=G)


class LM10Potential(MilkyWayPotential):
    """builtin/milkyway.py:101-169 (Law & Majewski 2010): MN disk, Hernquist bulge, triaxial logarithmic halo."""

    _defaults = {
        "disk": dict(m_tot=1e11, a=6.5, b=0.26),
        "bulge": dict(m_tot=3.4e10, r_s=0.7),
        "halo": dict(v_c=math.sqrt(2.0) * 121.858 * KMS, r_s=12.0, q1=1.38, q2=1.0, q3=1.36, phi=math.radians(97.0)),
    }
    _classes = {"disk": MiyamotoNagaiPotential, "bulge": HernquistPotential, "halo": LMJ09LogarithmicPotential}


class MilkyWayPotential2022(MilkyWayPotential):
    """builtin/milkyway.py:240-313: MN3Sech2 disk (positive density), NFW halo, two Hernquist spheres."""

    _defaults = {
        "disk": dict(m_tot=4.7717e10, h_R=2.6, h_z=0.3, positive_density=True),
        "halo": dict(m=5.5427e11, r_s=15.626),
        "bulge": dict(m_tot=5e9, r_s=1.0),
        "nucleus": dict(m_tot=1.8142e9, r_s=68.8867 * 0.001),
    }
    _classes = {"disk": MN3Sech2Potential, "halo": NFWPotential, "bulge": HernquistPotential,
                "nucleus": HernquistPotential}  # fmt: skip


class BovyMWPotential2014(MilkyWayPotential):
    """builtin/milkyway.py:34-97 (Bovy 2015 ``MWPotential2014``): MN disk, PowerLawCutoff bulge, NFW halo."""

    _defaults = {
        "disk": dict(m_tot=68_193_902_782.346756, a=3.0, b=280 * 0.001),
        "bulge": dict(m_tot=4501365375.06545, alpha=1.8, r_c=1.9),
        "halo": dict(m=4.3683325e11, r_s=16.0),
    }
    _classes = {"disk": MiyamotoNagaiPotential, "bulge": PowerLawCutoffPotential, "halo": NFWPotential}


__all__ = [
    "AbstractPotential", "LinearParameter", "ConstantParameter", "MiyamotoNagaiPotential", "HernquistPotential", "KeplerPotential", "PlummerPotential",
    "KuzminPotential", "IsochronePotential", "SatohPotential", "LogarithmicPotential", "LMJ09LogarithmicPotential",
    "LM10Potential", "NFWPotential", "TriaxialHernquistPotential", "JaffePotential", "BurkertPotential",
    "StoneOstriker15Potential", "HarmonicOscillatorPotential", "HenonHeilesPotential", "NullPotential",
    "PowerLawCutoffPotential", "MN3ExponentialPotential", "MN3Sech2Potential", "CompositePotential",
    "MilkyWayPotential", "MilkyWayPotential2022", "BovyMWPotential2014", "G_GALACTIC", "KMS",
]  # fmt: skip
