"""unxt stand-in: Quantity = (value, unit), ``ustrip(unit, q)``, unit systems as dicts of units.
Mirrors the calls the plugin makes: ``u.ustrip(unit, quantity)``, ``u.Q(value, unit)``, ``unit / unit``."""
import math

import numpy as np


class Unit:
    """A unit as its scale to the galactic system (kpc, Myr, Msun, rad)."""

    def __init__(self, name, scale):
        self.name, self.scale = name, float(scale)

    def __truediv__(self, other):
        return Unit(f"{self.name} / {other.name}", self.scale / other.scale)

    def __repr__(self):
        return f"Unit({self.name!r})"


KMS = 1.0 / 977.7922216807891  # km/s in kpc/Myr (what astropy's decomposition gives)
UNITS = {n: Unit(n, s) for n, s in {
    "kpc": 1.0, "pc": 1e-3, "Myr": 1.0, "Gyr": 1e3, "yr": 1e-6, "Msun": 1.0, "kpc / Myr": 1.0, "km / s": KMS, "rad": 1.0,
    "deg": math.pi / 180.0, "": 1.0, "kpc / Myr2": 1.0, "1 / Myr2": 1.0, "kpc2 / Myr2": 1.0, "Msun / Myr": 1.0,
    "Msun / yr": 1e6}.items()}  # fmt: skip


def unit(name):
    return name if isinstance(name, Unit) else UNITS[{"km/s": "km / s", "kpc/Myr": "kpc / Myr"}.get(name, name)]


class AbstractQuantity:
    pass


class Quantity(AbstractQuantity):
    def __init__(self, value, unit_):
        self.value, self.unit = np.asarray(value, dtype=np.float64), unit(unit_)

    @classmethod
    def from_(cls, value, unit_):
        return cls(value, unit_)

    @property
    def shape(self):
        return self.value.shape

    def __getitem__(self, i):
        return Quantity(self.value[i], self.unit)

    def __neg__(self):
        return Quantity(-self.value, self.unit)

    def __repr__(self):
        return f"Q({self.value!r}, {self.unit.name!r})"


Q = Quantity


def ustrip(*args):
    """``ustrip(unit, q)`` (and the ``ustrip(AllowValue, unit, x)`` form, where bare numbers pass through)."""
    if len(args) == 3:
        _, to, x = args
    else:
        to, x = args
    if isinstance(x, Quantity):
        return x.value * (x.unit.scale / unit(to).scale)
    return np.asarray(x, dtype=np.float64)


def uconvert(to, q):
    return Quantity(ustrip(to, q), to)


class AbstractUnitSystem(dict):
    pass


galactic = AbstractUnitSystem(
    length=UNITS["kpc"], mass=UNITS["Msun"], time=UNITS["Myr"], speed=UNITS["kpc / Myr"], angle=UNITS["rad"],
    dimensionless=UNITS[""], acceleration=UNITS["kpc / Myr2"], **{"frequency drift": UNITS["1 / Myr2"],
                                                                   "specific energy": UNITS["kpc2 / Myr2"]})  # fmt: skip
