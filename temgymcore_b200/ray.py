"""``Ray`` -- the SoA ray container (reference ``src/temgym_core/ray.py:8-166``).

Seven fp64 leaves ``x, y, dx, dy, z, pathlength, _one``; each may be a Python float or an
array (numpy array or torch tensor) of one common size.  This is exactly the layout the
CUDA ray kernel consumes and produces (one coalesced fp64 array per leaf).
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass
from typing import Any

import numpy as np

from . import _arrays as A
from .tree_utils import HasParamsMixin

RAY_FIELDS = ("x", "y", "dx", "dy", "z", "pathlength", "_one")


@dataclass(frozen=True, eq=False)
class Ray(HasParamsMixin):
    x: Any
    y: Any
    dx: Any
    dy: Any
    z: Any
    pathlength: Any
    _one: Any = 1.0

    @classmethod
    def origin(cls):  # ray.py:47-56
        return cls(*((0.0,) * 6))

    def _ray_items(self):
        return {f: getattr(self, f) for f in RAY_FIELDS}

    @property
    def size(self):  # ray.py:58-77
        sizes = set(A.numel(v) for v in self._ray_items().values())
        assert len(sizes) == 1
        return tuple(sizes)[0]

    def __getitem__(self, arg):  # ray.py:79-93
        return type(self)(**{k: v[arg] for k, v in self._ray_items().items()})

    def to_ray(self):
        return self

    def item(self):  # ray.py:98-117
        return type(self)(**{k: (v.item() if hasattr(v, "size") or hasattr(v, "numel") else v)
                             for k, v in self._ray_items().items()})

    def to_vector(self):  # ray.py:119-125
        def v1(v):
            if A.kind_of(v) in (A.KIND_TORCH_CPU, A.KIND_CUDA):
                return v.reshape(-1) if v.ndim == 0 else v
            return np.atleast_1d(np.asarray(v, dtype=np.float64))
        return type(self)(**{k: v1(v) for k, v in self._ray_items().items()})

    def derive(self, x=None, y=None, dx=None, dy=None, z=None, pathlength=None) -> "Ray":
        # ray.py:127-166 -- `_one` is preserved
        return Ray(
            x=x if x is not None else self.x,
            y=y if y is not None else self.y,
            dx=dx if dx is not None else self.dx,
            dy=dy if dy is not None else self.dy,
            z=z if z is not None else self.z,
            pathlength=pathlength if pathlength is not None else self.pathlength,
            _one=self._one,
        )
