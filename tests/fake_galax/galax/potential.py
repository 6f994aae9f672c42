"""galax.potential stand-in.  Attribute names follow the reference:
``pot.constants["G"].value`` (potential/_src/base.py:35,91-95), ``pot.units[...]``, parameters as
``ConstantParameter(value=Quantity)`` / ``LinearParameter(slope, point_time, point_value)``
(potential/_src/params/constant.py, params/core.py:25-110; lengths given in pc are converted at construction,
params/field.py:206-223), composites as ordered mappings (potential/_src/base_multi.py, composite.py:32-111),
default parameter values of the named models from builtin/milkyway.py:65-97,203-236,275-313."""
import unxt as u

G_GALACTIC = 4.498502151469553e-12  # astropy CODATA G in kpc^3 / (Msun Myr^2)


class ConstantParameter:
    def __init__(self, value):
        self.value = value


class LinearParameter:
    def __init__(self, slope, point_time, point_value):
        self.slope, self.point_time, self.point_value = slope, point_time, point_value


class UserParameter:  # a general function of time: must be refused
    def __init__(self, func):
        self.func = func


class _G:
    value = G_GALACTIC


class AbstractPotential:
    _fields: tuple = ()
    _dims: dict = {}

    def __init__(self, *, units=None, **params):
        self.units = units or u.galactic
        self.constants = {"G": _G()}
        for name in self._fields:
            v = params.pop(name)
            if isinstance(v, (ConstantParameter, LinearParameter, UserParameter)):
                setattr(self, name, v)
            else:
                q = v if isinstance(v, u.Quantity) else u.Q(v, self.units[self._dims[name]])
                setattr(self, name, ConstantParameter(u.uconvert(self.units[self._dims[name]], q)))
        for k, v in params.items():
            setattr(self, k, v)


def _kind(name, fields, dims):
    return type(name, (AbstractPotential,), {"_fields": fields, "_dims": dict(zip(fields, dims))})


MiyamotoNagaiPotential = _kind("MiyamotoNagaiPotential", ("m_tot", "a", "b"), ("mass", "length", "length"))
HernquistPotential = _kind("HernquistPotential", ("m_tot", "r_s"), ("mass", "length"))
KeplerPotential = _kind("KeplerPotential", ("m_tot",), ("mass",))
NFWPotential = _kind("NFWPotential", ("m", "r_s"), ("mass", "length"))
PowerLawCutoffPotential = _kind("PowerLawCutoffPotential", ("m_tot", "alpha", "r_c"), ("mass", "dimensionless", "length"))
PlummerPotential = _kind("PlummerPotential", ("m_tot", "r_s"), ("mass", "length"))
KuzminPotential = _kind("KuzminPotential", ("m_tot", "r_s"), ("mass", "length"))
IsochronePotential = _kind("IsochronePotential", ("m_tot", "r_s"), ("mass", "length"))
SatohPotential = _kind("SatohPotential", ("m_tot", "a", "b"), ("mass", "length", "length"))
JaffePotential = _kind("JaffePotential", ("m_tot", "r_s"), ("mass", "length"))
BurkertPotential = _kind("BurkertPotential", ("m", "r_s"), ("mass", "length"))
StoneOstriker15Potential = _kind("StoneOstriker15Potential", ("m_tot", "r_c", "r_h"), ("mass", "length", "length"))
TriaxialHernquistPotential = _kind("TriaxialHernquistPotential", ("m_tot", "r_s", "q1", "q2"),
                                   ("mass", "length", "dimensionless", "dimensionless"))
LMJ09LogarithmicPotential = _kind("LMJ09LogarithmicPotential", ("v_c", "r_s", "q1", "q2", "q3", "phi"),
                                  ("speed", "length", "dimensionless", "dimensionless", "dimensionless", "angle"))
LogarithmicPotential = _kind("LogarithmicPotential", ("v_c", "r_s"), ("speed", "length"))


class _MN3(AbstractPotential):
    _fields = ("m_tot", "h_R", "h_z")
    _dims = {"m_tot": "mass", "h_R": "length", "h_z": "length"}

    def __init__(self, *, positive_density=False, **kw):
        super().__init__(**kw)
        self.positive_density = positive_density


class MN3Sech2Potential(_MN3):
    pass


class MN3ExponentialPotential(_MN3):
    pass


class NullPotential(AbstractPotential):
    pass


class AbstractCompositePotential(AbstractPotential):
    def __init__(self, components: dict, units=None):
        super().__init__(units=units)
        self._data = dict(components)

    def items(self):
        return self._data.items()

    def values(self):
        return self._data.values()

    def __getitem__(self, k):
        return self._data[k]


class CompositePotential(AbstractCompositePotential):
    pass


class MilkyWayPotential(AbstractCompositePotential):  # builtin/milkyway.py:203-236
    def __init__(self, **over):
        d = dict(disk=MiyamotoNagaiPotential(m_tot=6.8e10, a=3.0, b=0.28), halo=NFWPotential(m=5.4e11, r_s=15.62),
                 bulge=HernquistPotential(m_tot=5e9, r_s=1.0), nucleus=HernquistPotential(m_tot=1.71e9, r_s=0.07))
        d.update(over)
        super().__init__(d)


class MilkyWayPotential2022(AbstractCompositePotential):  # builtin/milkyway.py:275-313 (nucleus r_s given in pc)
    def __init__(self, **over):
        d = dict(disk=MN3Sech2Potential(m_tot=4.7717e10, h_R=2.6, h_z=0.3, positive_density=True),
                 halo=NFWPotential(m=5.5427e11, r_s=15.626), bulge=HernquistPotential(m_tot=5e9, r_s=1.0),
                 nucleus=HernquistPotential(m_tot=1.8142e9, r_s=u.Q(68.8867, "pc")))
        d.update(over)
        super().__init__(d)


class BovyMWPotential2014(AbstractCompositePotential):  # builtin/milkyway.py:65-97 (disk b given in pc)
    def __init__(self, **over):
        d = dict(disk=MiyamotoNagaiPotential(m_tot=68_193_902_782.346756, a=3.0, b=u.Q(280, "pc")),
                 bulge=PowerLawCutoffPotential(m_tot=4501365375.06545, alpha=1.8, r_c=1.9),
                 halo=NFWPotential(m=4.3683325e11, r_s=16.0))
        d.update(over)
        super().__init__(d)


class LM10Potential(AbstractCompositePotential):  # builtin/milkyway.py:101-169
    def __init__(self, **over):
        import math

        d = dict(disk=MiyamotoNagaiPotential(m_tot=1e11, a=6.5, b=0.26), bulge=HernquistPotential(m_tot=3.4e10, r_s=0.7),
                 halo=LMJ09LogarithmicPotential(v_c=u.Q(math.sqrt(2.0) * 121.858, "km / s"), r_s=12.0, q1=1.38, q2=1.0,
                                                q3=1.36, phi=u.Q(97.0, "deg")))
        d.update(over)
        super().__init__(d)
