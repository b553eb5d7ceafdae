"""Fused 4D-STEM shadow-image backprojection (BASELINE config C5).

The reference ships the building blocks -- ``Scanner`` / ``Descanner`` / ``DescanError``
(components.py:27-115, 252-372), ``ScanGrid`` / ``Detector`` pixel<->metre maps (grid.py:120-182),
``transfer_rays_pt_src`` (transfer.py:57-123), ``inplace_sum`` (utils.py:83-114) -- while the
composite workflow lives in a private downstream repository.  This module assembles
those blocks: every (scan position, detector pixel) pair is a ray that is traced back from its
detector pixel to the sample plane through the system's ABCD matrices (descan error included as the
scan-position-dependent 5th column) and its intensity is accumulated on the sample grid.  The
per-ray work runs in one CUDA kernel (``tg_stem4d_backproject``, csrc/stem4d.cu); the handful of 5x5
system matrices come from the CUDA ray kernel (``run_to_end_abcd``).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _arrays as A
from . import _lib as L
from .ray import Ray
from .run import run_to_end_abcd


def _rows(T):
    """3x3 (y, x, 1) transform -> (y-form, x-form) coefficient rows [a, b, c]: value = (a*row + b*col) + c."""
    return [T[0, 0], T[0, 1], T[0, 2], T[1, 0], T[1, 1], T[1, 2]]


def system_geometry(model_fn, scan_grid, detector, source_xy=(0.0, 0.0), out_grid=None, abcd_fn=None):
    """The 42 geometry doubles + 6 shapes of ``tg_stem4d_backproject``.

    ``model_fn(spx, spy)`` returns the component list for a scan position (metres); the sample
    plane is where ``scan_grid`` sits in it.  The 4x4 ABCD blocks must not depend on the scan
    position and the 5th column must be affine in it (true for Scanner / Descanner); both are
    verified.  ``abcd_fn(ray, model) -> (5,5)`` defaults to the CUDA ray kernel.
    """
    out_grid = scan_grid if out_grid is None else out_grid
    if abcd_fn is None:
        abcd_fn = lambda ray, model: np.asarray(run_to_end_abcd(ray, model, want_rays=False)[1])  # noqa: E731
    r0 = np.array([float(source_xy[0]), float(source_xy[1])])

    def mats(spx, spy):
        model = list(model_fn(spx, spy))
        idx = next((i for i, c in enumerate(model) if c is scan_grid), None)
        if idx is None:
            raise ValueError("scan_grid is not an element of the model returned by model_fn")
        ray = Ray(x=r0[0], y=r0[1], dx=0.0, dy=0.0, z=float(model[0].z), pathlength=0.0)
        return abcd_fn(ray, model), abcd_fn(ray, model[:idx + 1])

    d00, s00 = mats(0.0, 0.0)
    d10, s10 = mats(1.0, 0.0)
    d01, s01 = mats(0.0, 1.0)
    d11, s11 = mats(1.0, 1.0)
    for a, b, c, d in ((d00, d10, d01, d11), (s00, s10, s01, s11)):
        scale = max(1.0, float(np.abs(a[:4, :4]).max()))
        if not (np.allclose(a[:4, :4], b[:4, :4], rtol=1e-12, atol=1e-12 * scale)
                and np.allclose(a[:4, :4], c[:4, :4], rtol=1e-12, atol=1e-12 * scale)
                and np.allclose(a[:, 4] + (b[:, 4] - a[:, 4]) + (c[:, 4] - a[:, 4]), d[:, 4], rtol=1e-9,
                                atol=1e-12 * max(1.0, float(np.abs(d[:, 4]).max())))):
            raise ValueError("the model is not affine in the scan position")
    Adet, Bdet = d00[0:2, 0:2], d00[0:2, 2:4]
    Asamp, Bsamp = s00[0:2, 0:2], s00[0:2, 2:4]
    geom = []
    geom += _rows(scan_grid.pixels_to_metres_mat)
    geom += _rows(detector.pixels_to_metres_mat)
    geom += _rows(out_grid.metres_to_pixels_mat)
    geom += list(Adet @ r0)
    geom += [d00[0, 4], d00[1, 4], d10[0, 4] - d00[0, 4], d10[1, 4] - d00[1, 4],
             d01[0, 4] - d00[0, 4], d01[1, 4] - d00[1, 4]]
    geom += list(np.linalg.inv(Bdet).reshape(-1))
    geom += list(Asamp @ r0)
    geom += list(Bsamp.reshape(-1))
    geom += [s00[0, 4], s00[1, 4], s10[0, 4] - s00[0, 4], s10[1, 4] - s00[1, 4],
             s01[0, 4] - s00[0, 4], s01[1, 4] - s00[1, 4]]
    shapes = [int(scan_grid.shape[0]), int(scan_grid.shape[1]), int(detector.shape[0]), int(detector.shape[1]),
              int(out_grid.shape[0]), int(out_grid.shape[1])]
    return shapes, [float(v) for v in geom]


def backproject_4dstem(data4d, model_fn, scan_grid, detector, *, source_xy=(0.0, 0.0), out_grid=None,
                       scan_range=None, out=None, geometry=None, stepwise_only=False, kernel="auto"):
    """Sum every detector pixel of every scan position onto the sample grid -> ``(Oy, Ox)`` float32.

    data4d: ``(Sy, Sx, Dy, Dx)`` float32 or uint16 (numpy / torch; stays on the GPU if it is
    there).  ``scan_range=(begin, count)`` restricts to a shard of flattened scan positions;
    ``out`` accumulates into an existing CUDA image.  ``kernel`` selects the implementation for
    A/B checks: "auto" (integer DDA kernels when a frame's footprint fits the 64 x 64 shared-memory
    tile -- the single-crossing variant when 7 |slope| < 1 px/px, else the run-merging one -- else the
    guarded fp64 affine kernel, else the step-wise kernel), "dda" (run-merging DDA only), "affine"
    (never a DDA kernel) or "stepwise" (= ``stepwise_only``).  All give identical pixel indices.
    """
    if kernel not in ("auto", "dda", "affine", "stepwise"):
        raise ValueError(f"unknown kernel {kernel!r}")
    mode_bits = 2 if (stepwise_only or kernel == "stepwise") else {"affine": 4, "dda": 8}.get(kernel, 0)
    import torch
    lib = L.load()
    shapes, geom = geometry if geometry is not None else system_geometry(model_fn, scan_grid, detector, source_xy,
                                                                          out_grid)
    kind = A.kind_of(data4d)
    dev = A.cuda_device_of((data4d,)) or torch.device("cuda", A.current_device_index())
    d = data4d if isinstance(data4d, torch.Tensor) else torch.as_tensor(np.asarray(data4d))
    if d.dtype not in (torch.float32, torch.uint16):
        d = d.to(torch.float32)
    d = d.to(dev).contiguous()
    nscan = shapes[0] * shapes[1]
    begin, count = (0, nscan) if scan_range is None else scan_range
    if scan_range is None:
        if tuple(d.shape) != tuple(shapes[:4]):
            raise ValueError(f"data4d has shape {tuple(d.shape)}, expected {tuple(shapes[:4])}")
        base = d
    else:  # a shard holds only its own scan positions
        if d.numel() != count * shapes[2] * shapes[3]:
            raise ValueError("data4d shard does not match scan_range")
        base = d
    img = torch.zeros((shapes[4], shapes[5]), dtype=torch.float32, device=dev) if out is None else out
    # the kernel indexes frames by absolute scan position: offset the base pointer for shards
    ptr = base.data_ptr() - begin * shapes[2] * shapes[3] * base.element_size()
    with torch.cuda.device(dev):
        L.check(lib.tg_stem4d_backproject((C.c_int * 6)(*shapes), L.dbl_array(geom), ptr,
                                          int(d.dtype == torch.float32) | mode_bits, begin,
                                          count, img.data_ptr(),
                                          A.current_stream_ptr(dev)), "tg_stem4d_backproject")
    if out is not None or kind == A.KIND_CUDA:
        return img
    return img.cpu() if kind == A.KIND_TORCH_CPU else img.cpu().numpy()


def backproject_indices(model_fn, scan_grid, detector, *, source_xy=(0.0, 0.0), out_grid=None, geometry=None):
    """``(Sy*Sx, Dy*Dx, 2)`` int32 sample-grid pixel (py, px) of every ray (parity checks)."""
    import torch
    lib = L.load()
    shapes, geom = geometry if geometry is not None else system_geometry(model_fn, scan_grid, detector, source_xy,
                                                                          out_grid)
    dev = torch.device("cuda", A.current_device_index())
    n = shapes[0] * shapes[1]
    idx = torch.empty((n, shapes[2] * shapes[3], 2), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.tg_stem4d_indices((C.c_int * 6)(*shapes), L.dbl_array(geom), 0, n, idx.data_ptr(),
                                      A.current_stream_ptr(dev)), "tg_stem4d_indices")
    return idx.cpu().numpy()
