"""diffrax stand-in: the solver / controller records galax passes around (fields and defaults of diffrax 0.7.0)."""
import dataclasses


@dataclasses.dataclass(frozen=True)
class Dopri8:
    scan_kind: str | None = None


@dataclasses.dataclass(frozen=True)
class Dopri5:
    scan_kind: str | None = None


@dataclasses.dataclass(frozen=True)
class SemiImplicitEuler:
    pass


@dataclasses.dataclass(frozen=True)
class LeapfrogMidpoint:
    pass


@dataclasses.dataclass(frozen=True)
class ConstantStepSize:
    pass


@dataclasses.dataclass(frozen=True)
class PIDController:
    rtol: float
    atol: float
    pcoeff: float = 0.0
    icoeff: float = 1.0
    dcoeff: float = 0.0
    dtmin: float | None = None
    dtmax: float | None = None
    force_dtmin: bool = True
    factormin: float = 0.2
    factormax: float = 10.0
    safety: float = 0.9
