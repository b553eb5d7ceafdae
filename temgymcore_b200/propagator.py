"""Propagators (reference ``src/temgym_core/propagator.py``).

``FreeSpaceParaxial`` (propagator.py:43-76) is the propagator the CUDA ray
kernel implements: ``x += dx*d, y += dy*d, z += d, pathlength += d``.
``FreeSpaceDirCosine`` (propagator.py:79-116) is not on the accelerated path
(the reference itself notes it is not integrated with the ABCD matrices).
"""
from typing import NamedTuple


class BasePropagator:
    def __call__(self, ray, distance):
        raise NotImplementedError

    def with_distance(self, distance) -> "Propagator":
        return Propagator(distance, self)


class Propagator(NamedTuple):
    distance: float
    propagator: BasePropagator

    def __call__(self, ray):
        return self.propagator(ray, self.distance)


class FreeSpaceParaxial(BasePropagator):
    @staticmethod
    def propagate(ray, distance):
        """Propagate ``ray`` by ``distance`` on the GPU (propagator.py:52-72)."""
        from .run import _propagate_only
        return _propagate_only(ray, distance)

    @classmethod
    def __call__(cls, ray, distance):
        return cls.propagate(ray, distance)
