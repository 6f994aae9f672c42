import numpy as np

import unxt as u


class AbstractPos3D:
    pass


class CartesianPos3D(AbstractPos3D):
    def __init__(self, xyz: u.Quantity):
        self.xyz = xyz

    @classmethod
    def from_(cls, value, unit):
        return cls(u.Q(value, unit))


class CartesianAcc3D:
    def __init__(self, xyz: u.Quantity):
        self.xyz = xyz

    @classmethod
    def from_(cls, value, unit):
        return cls(u.Q(np.asarray(value), unit))
