"""galax.coordinates stand-in: phase-space containers with ``_qp(units=)`` (coordinates/_src/base.py:348-386)."""
import unxt as u


class _Frames:
    simulation_frame = "SimulationFrame()"


frames = _Frames()


class AbstractPhaseSpaceObject:
    def __init__(self, q, p, t=None, frame=None):
        self.q, self.p, self.t, self.frame = q, p, t, frame or frames.simulation_frame

    def _qp(self, *, units):
        return u.uconvert(units["length"], self.q), u.uconvert(units["speed"], self.p)

    @property
    def ndim(self):
        return self.q.value.ndim - 1


class AbstractPhaseSpaceCoordinate(AbstractPhaseSpaceObject):
    pass


class PhaseSpaceCoordinate(AbstractPhaseSpaceCoordinate):
    pass


class PhaseSpacePosition(AbstractPhaseSpaceObject):
    def __init__(self, q, p, frame=None):
        super().__init__(q, p, None, frame)
