"""``custom_jacobian_matrix`` (reference ``src/temgym_core/utils.py:7-43``) and the
Ray-of-Ray Jacobian container the CUDA kernel's 7x7 output is exposed through.

The reference gets a nested ``Ray``-of-``Ray`` pytree from ``jax.jacobian`` and picks the
``[x, y, dx, dy, _one]`` block out of it.  Here the kernel writes the Jacobian directly:
either already as the 5x5 ABCD block (``run_to_end_abcd``) or as the full 7x7
(``ray_jacobian``), which :class:`RayJacobian` presents with the same attribute access
(``jac.dx.x`` == d out.dx / d in.x).
"""
from __future__ import annotations

from .ray import RAY_FIELDS

_PICK = (0, 1, 2, 3, 6)


class _JacRow:
    def __init__(self, row):
        self._row = row

    def __getattr__(self, name):
        if name in RAY_FIELDS:
            return self._row[..., RAY_FIELDS.index(name)]
        raise AttributeError(name)


class RayJacobian:
    """``(..., 7, 7)`` array ``J[..., i, j] = d out_i / d in_j`` in ``RAY_FIELDS`` order."""

    def __init__(self, matrix):
        self.matrix = matrix

    def __getattr__(self, name):
        if name in RAY_FIELDS:
            return _JacRow(self.matrix[..., RAY_FIELDS.index(name), :])
        raise AttributeError(name)


class RayDerivative:
    """Order-k derivative of the output ray w.r.t. the input ray: ``tensor[..., f, a_1, ..., a_k] =
    d^k out_f / d in_a1 ... d in_ak`` in ``RAY_FIELDS`` order.  Attribute access peels one index at a
    time like the reference's nested Ray pytrees: ``derivs[1].x.dx.dy`` (run.py:119-147)."""

    def __init__(self, tensor, order, _depth=0):
        self.tensor = tensor
        self.order = order
        self._depth = _depth

    def __getattr__(self, name):
        if name in RAY_FIELDS:
            axis = self.tensor.ndim - (self.order + 1 - self._depth)
            sub = self.tensor[(slice(None),) * axis + (RAY_FIELDS.index(name),)]
            if self._depth == self.order:
                return sub
            return RayDerivative(sub, self.order, self._depth + 1)
        raise AttributeError(name)


def custom_jacobian_matrix(ray_jac):
    """-> ``(..., 5, 5)`` over ``[x, y, dx, dy, _one]`` (utils.py:34-43)."""
    m = ray_jac.matrix if isinstance(ray_jac, RayJacobian) else ray_jac
    if m.shape[-1] == 5 and m.shape[-2] == 5:
        return m
    if m.shape[-1] != 7 or m.shape[-2] != 7:
        raise ValueError(f"expected a (...,7,7) or (...,5,5) Jacobian, got {tuple(m.shape)}")
    idx = list(_PICK)
    return m[..., idx, :][..., :, idx]


def fibonacci_spiral(nb_samples: int, radius: float, alpha=2, device=None):
    """Beamlet-centre sampler (utils.py:297-325).  Host numpy by default like the reference; with
    ``device="cuda"`` (or a torch device) the points are generated on the GPU
    (``tg_fibonacci_spiral_f64``) and returned as CUDA tensors -- no host staging."""
    import numpy as np
    if device is not None:
        import torch
        from . import _arrays as A
        from . import _lib as L
        dev = torch.device(device)
        if dev.type != "cuda":
            raise ValueError("device must be a CUDA device (omit it for the host sampler)")
        if dev.index is None:
            dev = torch.device("cuda", A.current_device_index())
        x = torch.empty(int(nb_samples), dtype=torch.float64, device=dev)
        y = torch.empty_like(x)
        with torch.cuda.device(dev):
            L.check(L.load().tg_fibonacci_spiral_f64(int(nb_samples), float(radius), float(alpha), x.data_ptr(),
                                                     y.data_ptr(), A.current_stream_ptr(dev)),
                    "tg_fibonacci_spiral_f64")
        return x, y
    ga = np.pi * (3.0 - np.sqrt(5.0))
    np_boundary = np.round(alpha * np.sqrt(nb_samples))
    ii = np.arange(nb_samples)
    with np.errstate(invalid="ignore"):
        rr = np.where(ii > nb_samples - (np_boundary + 1), radius,
                      radius * np.sqrt((ii + 0.5) / (nb_samples - 0.5 * (np_boundary + 1))))
    rr[0] = 0.
    phi = ii * ga
    return rr * np.cos(phi), rr * np.sin(phi)


def concentric_rings(num_points_approx: int, radius: float, device=None):
    """Approximately uniform ``(y, x)`` samples on concentric rings of a disc
    (utils.py:117-175).  Ring k holds ~2*pi*k points; the angles of a ring start at 0 and advance by running
    sums of 2*pi/n_k, like the reference's ``multi_cumsum_inplace`` (utils.py:46-80).  Host numpy by default
    like the reference; with ``device="cuda"`` (or a torch device) the points are generated on the GPU
    (``tg_concentric_rings_f64``: same ring layout, same running sums) and returned as an ``(N, 2)`` CUDA
    tensor -- no host staging for 1e6-ray runs."""
    if device is not None:
        import torch
        from . import _arrays as A
        from . import _lib as L
        dev = torch.device(device)
        if dev.type != "cuda":
            raise ValueError("device must be a CUDA device (omit it for the host sampler)")
        if dev.index is None:
            dev = torch.device("cuda", A.current_device_index())
        lib = L.load()
        n = int(lib.tg_concentric_rings_count(int(num_points_approx), float(radius)))
        if n < 0:
            raise ValueError("bad point count")
        yx = torch.empty((2, max(n, 1)), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            L.check(lib.tg_concentric_rings_f64(int(num_points_approx), float(radius), n, yx[0].data_ptr(),
                                                yx[1].data_ptr(), A.current_stream_ptr(dev)),
                    "tg_concentric_rings_f64")
        return yx[:, :n].T
    import numpy as np
    n_rings = max(1, int(np.floor((-1 + np.sqrt(1 + 4 * num_points_approx / np.pi)) / 2)))
    circumference = np.round(2 * np.pi * np.arange(1, n_rings + 1)).astype(int)
    per_ring = np.round(circumference * (num_points_approx / circumference.sum())).astype(int)
    radii = np.linspace(0, radius, n_rings + 1, endpoint=True)[1:]
    reps = per_ring.tolist()
    all_radii = np.repeat(radii, reps)
    ang = np.repeat(2 * np.pi / per_ring, reps)
    # Running angle sums restart from 0 where the reference's multi_cumsum_inplace restarts them:
    # its partition counter lags by one, so segment k begins k elements after ring k does
    # (utils.py:69-80).  Reproduced as is: these are the ray directions users get.
    start, k = 0, 0
    while start < ang.size:
        stop = min(ang.size, start + reps[k] + 1)
        ang[start] = 0.0
        ang[start:stop] = np.cumsum(ang[start:stop])
        start, k = stop, min(k + 1, len(reps) - 1)
    return np.stack((all_radii * np.sin(ang), all_radii * np.cos(ang)), axis=-1)


def random_coords(num: int):
    """Uniform random ``(y, x)`` points in the unit disc, at least one (utils.py:178-205)."""
    import numpy as np
    while True:
        yx = np.random.uniform(-1, 1, size=(max(1, int(num * 1.28)), 2))
        keep = np.sqrt((yx ** 2).sum(axis=1)) < 1
        if keep.sum() > 0:
            return yx[keep, :]


# ------------------------------------------------------------------------------------------
# Host-side helpers of the reference's utils.py that sit beside the hot path (numpy, like the
# reference; ``xp`` may be any numpy-compatible namespace, e.g. ``torch`` is NOT one -- pass arrays
# of the namespace you choose).


def multi_cumsum_inplace(values, partitions, start):
    """Cumulative sums restarted per partition, in place (utils.py:46-80) -- including the reference's
    restart rule (the running counter is compared BEFORE it is advanced, so partition k restarts k
    elements late); ``concentric_rings`` depends on exactly this."""
    part_idx, part_count = 0, 0
    current = partitions[0]
    values[0] = start
    for i in range(1, len(values)):
        if current == part_count:
            part_count = 0
            part_idx += 1
            current = partitions[part_idx]
            values[i] = start
        else:
            values[i] += values[i - 1]
            part_count += 1


def inplace_sum(px_y, px_x, mask, frame, buffer):
    """``buffer[py, px] += frame`` for masked, in-bounds entries (utils.py:83-114), host arrays.
    (The CUDA histogram ``Grid.into_image`` / ``tg_into_image_i64`` is the device counterpart.)"""
    import numpy as np
    py, px = np.asarray(px_y), np.asarray(px_x)
    h, w = buffer.shape
    ok = np.asarray(mask, dtype=bool) & (py >= 0) & (py < h) & (px >= 0) & (px < w)
    np.add.at(buffer, (py[ok], px[ok]), np.asarray(frame)[ok])


def try_ravel(val):
    """``val.ravel()`` when it has one, else ``val`` (utils.py:208-225)."""
    try:
        return val.ravel()
    except AttributeError:
        return val


def try_reshape(val, maybe_has_shape):
    """``val.reshape(maybe_has_shape.shape)`` when possible, else ``val`` (utils.py:228-245)."""
    try:
        return val.reshape(maybe_has_shape.shape)
    except AttributeError:
        return val


def FresnelPropagator(u1, L, wavelength, z, xp=None):
    """Paraxial free-space propagation of a sampled field over ``z`` by the transfer-function method
    (utils.py:248-265): the validator the reference's wave-optics tests compare the beamlet sum with.
    ``L`` is the side length of the (square-pixel) window; the pixel pitch is ``L / rows``."""
    if xp is None and type(u1).__module__.startswith("torch"):
        # device-side validator: the same transfer-function method with torch.fft (cuFFT on
        # a CUDA tensor), so a beamlet image can be cross-checked against wave optics without leaving the GPU
        import math
        import torch
        rows, cols = u1.shape
        pitch = L / rows
        rdt = torch.float64 if u1.dtype in (torch.complex128, torch.float64) else torch.float32
        fx = torch.fft.fftfreq(cols, d=pitch, dtype=rdt, device=u1.device)
        fy = torch.fft.fftfreq(rows, d=pitch, dtype=rdt, device=u1.device)
        FY, FX = torch.meshgrid(fy, fx, indexing="ij")
        H = torch.exp(-1j * math.pi * wavelength * z * (FX ** 2 + FY ** 2))
        return torch.fft.ifft2(H * torch.fft.fft2(u1))
    if xp is None:
        import numpy as xp
    rows, cols = u1.shape
    pitch = L / rows
    fx = xp.fft.fftfreq(cols, d=pitch)
    fy = xp.fft.fftfreq(rows, d=pitch)
    FX, FY = xp.meshgrid(fx, fy)
    H = xp.exp(-1j * xp.pi * wavelength * z * (FX ** 2 + FY ** 2))
    return xp.fft.ifft2(H * xp.fft.fft2(u1))


def fresnel_lens_imaging_solution(E0, Y, X, ps, lambda0, z1, f, z2):
    """Fresnel propagation over z1, thin-lens phase, Fresnel propagation over z2 (utils.py:268-275)."""
    import numpy as np
    k = 2 * np.pi / lambda0
    L = E0.shape[0] * ps
    if type(E0).__module__.startswith("torch"):     # torch tensors (CPU or CUDA): stay on their device
        import torch
        X, Y = torch.as_tensor(X, device=E0.device), torch.as_tensor(Y, device=E0.device)
        lens = torch.exp((-1j * k) / (2 * f) * (X ** 2 + Y ** 2))
    else:
        lens = np.exp((-1j * k) / (2 * f) * (X ** 2 + Y ** 2))
    at_lens = FresnelPropagator(E0, L, lambda0, z1) * lens
    return FresnelPropagator(at_lens, L, lambda0, z2)


def zero_phase(u, idx_x, idx_y):
    """Rotate the global phase so that ``u[idx_x, idx_y]`` is real and positive, in place
    (utils.py:278-282)."""
    import numpy as np
    u *= np.exp(-1j * np.angle(u[idx_x, idx_y]))
    return u


def make_aperture(X, Y, aperture_ratio=0.1):
    """Boolean disc mask of radius ``aperture_ratio * max|X|`` (utils.py:285-294)."""
    import numpy as np
    return X ** 2 + Y ** 2 < (np.max(np.abs(X)) * aperture_ratio) ** 2
