"""Sources as model elements (reference ``src/temgym_core/source.py``).

In a model a ``Source`` is a no-op plane at its ``z`` (source.py:21-34): the ray kernel
only needs that ``z``.  Ray *generation* (``make_rays`` / ``generate_array``,
source.py:36-188) is host-side numpy input preparation in the reference; it is provided here
(numpy) so reference scripts run unchanged.  ``make_rays(num, device="cuda")`` generates the
deterministic ring pattern ON the device instead (``tg_concentric_rings_f64``) and returns a
``Ray`` of CUDA tensors that feeds ``run_to_end`` without host staging.
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

    def generate_array(self, num: int, random: bool = False):
        """(N, 5) rows ``[x, y, dx, dy, 1]`` (source.py:36-56)."""
        raise NotImplementedError

    def make_rays(self, num: int, random: bool = False, device=None):
        """``Ray`` with vector ``x, y, dx, dy`` and scalar ``z = self.z``, ``pathlength = 0``
        (source.py:58-79); a single generated ray gives scalar fields.  ``device``: generate on that CUDA
        device (deterministic mode: ring kernel; random mode draws with numpy's generator, like the
        reference, and uploads)."""
        from .ray import Ray
        if device is not None:
            x, y, dx, dy = self._generate_device(num, random, device)
            return Ray(x=x, y=y, dx=dx, dy=dy, z=self.z, pathlength=0.)
        r = self.generate_array(num, random=random)
        sl = 0 if r.shape[0] == 1 else slice(None)
        return Ray(x=r[sl, 0], y=r[sl, 1], dx=r[sl, 2], dy=r[sl, 3], z=self.z, pathlength=0.)

    def _disc(self, num, scale, random, device=None):
        from .utils import concentric_rings, random_coords
        if device is None:
            return random_coords(num) * scale if random else concentric_rings(num, scale)
        import torch
        if random:
            return torch.as_tensor(random_coords(num) * scale, device=device)
        return concentric_rings(num, scale, device=device)

    def _generate_device(self, num, random, device):
        raise NotImplementedError


@dataclass(frozen=True)
class PointSource(Source):
    z: float
    semi_conv: float
    offset_xy: Any = (0.0, 0.0)

    def generate_array(self, num: int, random: bool = False):
        # all rays leave the offset point; slopes fill the cone of semi-convergence (source.py:107-135)
        import numpy as np
        dy, dx = self._disc(num, self.semi_conv, random).T
        r = np.zeros((dx.size, 5), dtype=np.float64)
        r[:, 0] += self.offset_xy[0]
        r[:, 1] += self.offset_xy[1]
        r[:, 2], r[:, 3], r[:, 4] = dx, dy, 1.0
        return r

    def _generate_device(self, num, random, device):
        import torch
        dyx = self._disc(num, self.semi_conv, random, device)
        dy, dx = dyx[:, 0].contiguous(), dyx[:, 1].contiguous()
        x = torch.zeros_like(dx) + float(self.offset_xy[0])
        y = torch.zeros_like(dx) + float(self.offset_xy[1])
        return x, y, dx, dy


@dataclass(frozen=True)
class ParallelBeam(Source):
    z: float
    radius: float
    offset_xy: Any = (0.0, 0.0)

    def generate_array(self, num: int, random: bool = False):
        # parallel rays (zero slope) filling the aperture disc (source.py:160-188)
        import numpy as np
        y, x = self._disc(num, self.radius, random).T
        r = np.zeros((x.size, 5), dtype=np.float64)
        r[:, 0], r[:, 1], r[:, 4] = x + self.offset_xy[0], y + self.offset_xy[1], 1.0
        return r

    def _generate_device(self, num, random, device):
        import torch
        yx = self._disc(num, self.radius, random, device)
        x = yx[:, 1] + float(self.offset_xy[0])
        y = yx[:, 0] + float(self.offset_xy[1])
        return x.contiguous(), y.contiguous(), torch.zeros_like(x), torch.zeros_like(x)
