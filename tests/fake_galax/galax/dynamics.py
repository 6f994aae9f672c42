"""galax.dynamics stand-in: the containers and solver records the plugin reads and returns
(dynamics/_src/orbit/orbit.py:18-75, legacy/integrator.py:42-245, orbit/solver.py:121-141,
legacy/mockstream/{core.py,mockstream_generator.py:40-60}, df/fardal15.py, df/chen24.py)."""
import dataclasses

import diffrax as dfx


@dataclasses.dataclass
class Orbit:
    q: object
    p: object
    t: object
    frame: object = None


@dataclasses.dataclass(frozen=True)
class OrbitSolver:
    solver: object = dataclasses.field(default_factory=dfx.Dopri8)
    stepsize_controller: object = dataclasses.field(default_factory=lambda: dfx.PIDController(rtol=1e-8, atol=1e-8))
    max_steps: int | None = 2**16


@dataclasses.dataclass(frozen=True)
class Integrator:
    dynamics_solver: OrbitSolver = dataclasses.field(
        default_factory=lambda: OrbitSolver(stepsize_controller=dfx.PIDController(rtol=1e-7, atol=1e-7)))
    diffeq_kw: dict = dataclasses.field(default_factory=lambda: {"max_steps": None})


class _Fields:
    class HamiltonianField:
        def __init__(self, potential):
            self.potential = potential

        @property
        def units(self):
            return self.potential.units


fields = _Fields()


@dataclasses.dataclass
class MockStreamArm:
    q: object
    p: object
    t: object
    release_time: object
    frame: object = None


class MockStream(dict):
    pass


class FardalStreamDF:
    pass


class ChenStreamDF:
    pass
