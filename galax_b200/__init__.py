"""galax_b200 -- B200-native (sm_100a) implementation of galax's data-parallel hot path.

Batched fp64 test-particle orbit integration in analytic composite Milky-Way potentials, the mock-stream
path built on it, and bulk potential / acceleration / Hessian evaluation, behind a mirror of the
``galax.potential`` / ``galax.dynamics`` API.  All numerics run in ``libgalax_b200.so`` (hand-written CUDA,
C ABI in ``include/galax_b200.h``); there is no CPU fallback.

    import galax_b200.potential as gp
    import galax_b200.dynamics as gd
"""

from . import _lib, dynamics, experimental, potential
from ._lib import GalaxB200Error, build

__all__ = ["potential", "dynamics", "experimental", "build", "GalaxB200Error"]
__version__ = "0.1.0"
