"""``Grid`` mixin: pixel <-> metre conversions (reference ``src/temgym_core/grid.py``).

The 3x3 matrices are built on the host in the reference's composition order
(coordinate_transforms.py:92-99) and inverted with LU (``np.linalg.inv``, as
``jnp.linalg.inv`` at grid.py:59-63); the per-point work -- the 3-term affine form,
``round`` half-to-even and the int32 cast (grid.py:142-153) -- runs on the GPU in
``tg_metres_to_pixels`` with a fixed, FMA-free evaluation order so pixel indices are
bit-reproducible.  ``into_image`` (grid.py:231-280; numba ``inplace_sum``, utils.py:83-114)
is an ``atomicAdd`` histogram kernel.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import CoordsXY, PixelsYX
from . import _arrays as A
from . import _lib as L
from .coordinate_transforms import pixels_to_metres_transform


def _affine_apply(y, x, M: np.ndarray, as_float: bool):
    """(T @ [y, x, 1])[:2] per point on the GPU; int32 (rounded) unless ``as_float``."""
    lib = L.load()
    ky, kx = A.kind_of(y), A.kind_of(x)
    kind = max(ky, kx)
    n = max(A.numel(y), A.numel(x))
    out_shape = A.shape_of(y) if ky != A.KIND_SCALAR else A.shape_of(x)
    m9 = L.dbl_array(np.asarray(M, dtype=np.float64).reshape(-1))
    if kind == A.KIND_CUDA:
        import torch
        dev = A.cuda_device_of((y, x))
        ty = A.to_device_f64(y, dev).expand(n).contiguous() if A.numel(y) != n else A.to_device_f64(y, dev)
        tx = A.to_device_f64(x, dev).expand(n).contiguous() if A.numel(x) != n else A.to_device_f64(x, dev)
        dt = torch.float64 if as_float else torch.int32
        oy = torch.empty(n, dtype=dt, device=dev)
        ox = torch.empty(n, dtype=dt, device=dev)
        with torch.cuda.device(dev):
            L.check(lib.tg_metres_to_pixels(n, tx.data_ptr(), ty.data_ptr(), m9, oy.data_ptr(),
                                            ox.data_ptr(), int(as_float), A.current_stream_ptr(dev)),
                    "tg_metres_to_pixels")
        return oy.reshape(out_shape), ox.reshape(out_shape)
    hy = np.broadcast_to(A.to_host_f64(y), (n,)) if A.numel(y) != n else A.to_host_f64(y)
    hx = np.broadcast_to(A.to_host_f64(x), (n,)) if A.numel(x) != n else A.to_host_f64(x)
    hy, hx = np.ascontiguousarray(hy), np.ascontiguousarray(hx)
    dt = np.float64 if as_float else np.int32
    oy, ox = np.empty(n, dtype=dt), np.empty(n, dtype=dt)
    L.check(lib.tg_metres_to_pixels_host(n, hx.ctypes.data, hy.ctypes.data, m9, oy.ctypes.data,
                                         ox.ctypes.data, int(as_float), A.current_device_index()),
            "tg_metres_to_pixels_host")
    if kind == A.KIND_SCALAR:
        return oy.reshape(())[()], ox.reshape(())[()]
    return A.from_host(oy, kind, out_shape), A.from_host(ox, kind, out_shape)


_PX2M_MEMO: dict = {}


class Grid:
    """Mixin for components with ``z, centre, shape, pixel_size, rotation, flip_y``."""

    @property
    def pixels_to_metres_mat(self) -> np.ndarray:  # grid.py:37-48
        return pixels_to_metres_transform(self.centre, self.pixel_size, self.shape, self.flip_y,
                                          self.rotation)

    @property
    def metres_to_pixels_mat(self) -> np.ndarray:  # grid.py:50-63
        return np.linalg.inv(self.pixels_to_metres_mat)

    @property
    def px2m_affine(self):
        """``(X0, Xc, Xr, Y0, Yc, Yr)``: x_m = X0 + Xc*col + Xr*row, y_m likewise --
        the six doubles the field-sum kernel regenerates ``coords`` from.  Memoised on the grid's
        parameters: building the 3x3 matrix costs ~30 us of numpy, a tenth of a whole C2 image call."""
        try:
            key = (tuple(float(v) for v in self.centre), tuple(float(v) for v in self.pixel_size),
                   tuple(int(v) for v in self.shape), bool(self.flip_y), float(self.rotation))
        except (TypeError, ValueError):       # array-valued parameters: no memo
            key = None
        hit = _PX2M_MEMO.get(key) if key is not None else None
        if hit is None:
            T = self.pixels_to_metres_mat
            hit = (float(T[1, 2]), float(T[1, 1]), float(T[1, 0]), float(T[0, 2]), float(T[0, 1]), float(T[0, 0]))
            if key is not None:
                if len(_PX2M_MEMO) > 256:
                    _PX2M_MEMO.clear()
                _PX2M_MEMO[key] = hit
        return hit

    @property
    def coords_px(self) -> PixelsYX:  # grid.py:65-80
        yy, xx = np.meshgrid(np.arange(self.shape[0]), np.arange(self.shape[1]), indexing="ij")
        return PixelsYX(yy, xx)

    @property
    def coords(self) -> np.ndarray:  # grid.py:82-100 -> (H*W, 2) of (x_m, y_m), row-major
        yy, xx = self.coords_px
        cx, cy = self.pixels_to_metres((yy.ravel(), xx.ravel()))
        return np.stack((cx, cy), axis=-1).reshape(-1, 2)

    @property
    def coords_1d(self):  # grid.py:102-118
        yy, xx = self.coords_px
        x_coords = self.pixels_to_metres((yy[0, :], xx[0, :]))[0]
        y_coords = self.pixels_to_metres((yy[:, 0], xx[:, 0]))[1]
        return x_coords, y_coords

    def metres_to_pixels(self, coords, cast: bool = True) -> PixelsYX:  # grid.py:120-153
        coords_x, coords_y = coords
        py, px = _affine_apply(coords_y, coords_x, self.metres_to_pixels_mat, as_float=not cast)
        return PixelsYX(y=py, x=px)

    def pixels_to_metres(self, pixels) -> CoordsXY:  # grid.py:155-182
        pixels_y, pixels_x = pixels
        my, mx = _affine_apply(pixels_y, pixels_x, self.pixels_to_metres_mat, as_float=True)
        return CoordsXY(x=mx, y=my)

    def ray_at_grid(self, px_y, px_x, dx=0., dy=0., z=None):  # grid.py:184-212
        from .ray import Ray
        if z is None:
            z = self.z
        x, y = self.pixels_to_metres((px_y, px_x))
        return Ray(x=x, y=y, dx=dx, dy=dy, z=z, pathlength=0.)

    def ray_to_grid(self, ray, cast: bool = False) -> PixelsYX:  # grid.py:214-229
        return self.metres_to_pixels((ray.x, ray.y), cast=cast)

    def into_image(self, ray, acc=None):  # grid.py:231-280
        from .ray import Ray
        lib = L.load()
        if isinstance(ray, Ray):
            yy, xx = self.ray_to_grid(ray, cast=True)
        else:
            yy, xx = ray
        H, W = int(self.shape[0]), int(self.shape[1])
        if acc is not None and tuple(acc.shape) != (H, W):
            raise ValueError("acc has a different shape than the grid")
        import torch
        kind = max(A.kind_of(yy), A.kind_of(xx))
        dev = A.cuda_device_of((yy, xx)) or torch.device("cuda", A.current_device_index())
        ty = torch.as_tensor(np.asarray(yy) if kind < A.KIND_TORCH_CPU else yy).to(dev, torch.int32).reshape(-1).contiguous()
        tx = torch.as_tensor(np.asarray(xx) if kind < A.KIND_TORCH_CPU else xx).to(dev, torch.int32).reshape(-1).contiguous()
        img = torch.zeros((H, W), dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            L.check(lib.tg_into_image_i64(ty.numel(), ty.data_ptr(), tx.data_ptr(), H, W, img.data_ptr(),
                                          A.current_stream_ptr(dev)), "tg_into_image_i64")
        if kind == A.KIND_CUDA:
            if acc is not None:
                acc += img.to(acc.dtype)
                return acc
            return img
        res = img.cpu().numpy()
        if acc is None:
            return res.astype(int)
        acc += res.astype(acc.dtype)
        return acc
