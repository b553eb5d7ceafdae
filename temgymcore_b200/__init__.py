"""temgymcore_b200 -- B200-native (sm_100a) implementation of TemGymCore's hot path.

Mirrors the reference's call surface (module and symbol names of
``src/temgym_core``) for the one data-parallel path it accelerates:
batched ray propagation + 5x5 ABCD matrices, and the Gaussian-beamlet field
sum.  All compute runs in hand-written CUDA kernels behind the C ABI declared
in ``include/temgym_b200.h`` (``libtemgym_b200.so``); there is no CPU fallback.

Type aliases / NamedTuples follow reference ``src/temgym_core/__init__.py:10-125``.
"""
from typing import NamedTuple, Union, Any

__version__ = "0.1.0"

PositiveFloat = float
NonNegativeFloat = float
Radians = float
Degrees = float


class ShapeYX(NamedTuple):
    y: int
    x: int


class ScaleYX(NamedTuple):
    y: float
    x: float


class CoordXY(NamedTuple):
    x: float
    y: float

    def to_coords(self) -> "CoordsXY":
        import numpy as np
        return CoordsXY(x=np.array((self.x,)), y=np.array((self.y,)))


class CoordsXY(NamedTuple):
    x: Any
    y: Any


class PixelYX(NamedTuple):
    y: Union[int, float]
    x: Union[int, float]

    def to_pixels(self) -> "PixelsYX":
        import numpy as np
        return PixelsYX(x=np.array((self.x,)), y=np.array((self.y,)))


class PixelsYX(NamedTuple):
    y: Any
    x: Any
