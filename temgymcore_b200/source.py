"""Sources as model elements (reference ``src/temgym_core/source.py``).

In a model a ``Source`` is a no-op plane at its ``z`` (source.py:21-34): the ray kernel
only needs that ``z``.  Ray *generation* (``make_rays`` / ``generate_array``,
source.py:36-188) is host-side numpy input preparation and outside the accelerated path
(SURVEY.md section 2); ``ParallelBeam`` / ``PointSource`` are provided so reference models
can be written unchanged.
"""
from dataclasses import dataclass
from typing import Any

from .tree_utils import HasParamsMixin


class Source(HasParamsMixin):
    z: float

    def __call__(self, ray):
        return ray

    def _tg_param_seeds(self, path):
        return [(0, 1.0)] if tuple(path) == ("z",) else []


@dataclass(frozen=True)
class PointSource(Source):
    z: float
    semi_conv: float
    offset_xy: Any = (0.0, 0.0)


@dataclass(frozen=True)
class ParallelBeam(Source):
    z: float
    radius: float
    offset_xy: Any = (0.0, 0.0)
