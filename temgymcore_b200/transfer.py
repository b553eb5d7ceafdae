"""5x5 transfer-matrix helpers (reference ``src/temgym_core/transfer.py``).

The matrix products are tiny host algebra (numpy, as in the reference); applying the matrices
to a batch of rays -- ``transfer_rays`` / ``transfer_rays_pt_src`` -- runs on the GPU
(``tg_transfer_rays_f64``, csrc/trace.cu).
"""
from __future__ import annotations

import numpy as np

from . import _arrays as A
from . import _lib as L


def _np(v):
    return v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)


def accumulate_matrices(matrices):
    """``matrices[-1] @ ... @ matrices[0]`` (transfer.py:126-147)."""
    total = _np(matrices[-1])
    for tm in reversed(matrices[:-1]):
        total = total @ _np(tm)
    return total


def accumulate_matrices_cumulative(matrices):
    """Cumulative products in the reference's order (transfer.py:150-183): entry 0 is the LAST
    matrix, entry k is ``matrices[-1-k] @ ... @ matrices[-1]`` -- the loop of the reference,
    kept as it is (its docstring describes the opposite order)."""
    total = _np(matrices[-1])
    out = [total]
    for tm in reversed(matrices[:-1]):
        total = _np(tm) @ total
        out.append(total)
    return np.stack(out, axis=0)


def _apply(rays, mats):
    """out[n, m, :] = mats[m] @ rays[n] on the GPU; rays (N,5), mats (M,5,5) host."""
    import torch
    lib = L.load()
    kind = A.kind_of(rays)
    dev = A.cuda_device_of((rays,)) or torch.device("cuda", A.current_device_index())
    r = rays.detach().to(dev, torch.float64) if kind >= A.KIND_TORCH_CPU else \
        torch.as_tensor(np.asarray(rays, dtype=np.float64), device=dev)
    r = r.reshape(-1, 5).contiguous()
    mats = np.ascontiguousarray(np.asarray(mats, dtype=np.float64).reshape(-1, 5, 5))
    n, m = r.shape[0], mats.shape[0]
    out = torch.empty((n, m, 5), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        for m0 in range(0, m, 32):
            mm = min(32, m - m0)
            part = out if (m0 == 0 and mm == m) else torch.empty((n, mm, 5), dtype=torch.float64, device=dev)
            L.check(lib.tg_transfer_rays_f64(n, r.data_ptr(), mm,
                                             L.dbl_array(mats[m0:m0 + mm].reshape(-1)), part.data_ptr(),
                                             A.current_stream_ptr(dev)), "tg_transfer_rays_f64")
            if part is not out:
                out[:, m0:m0 + mm] = part
    if kind == A.KIND_CUDA:
        return out
    return out.cpu() if kind == A.KIND_TORCH_CPU else out.cpu().numpy()


def transfer_rays(ray_coords, transfer_matrices):
    """Apply the cumulative 5x5 matrices to a batch of rays -> ``(N, M, 5)``
    (transfer.py:6-54)."""
    if len(transfer_matrices) == 0:
        raise IndexError("transfer_matrices is empty")
    cumulative = accumulate_matrices_cumulative(transfer_matrices)
    return _apply(ray_coords, cumulative)


def transfer_rays_pt_src(input_pos_xy, input_slopes_xy, transfer_matrix):
    """Rays from a point source through one 5x5 matrix -> ``(4, N)`` rows ``[x, y, dx, dy]``
    (transfer.py:57-123)."""
    x0, y0 = input_pos_xy
    sx, sy = input_slopes_xy
    kind = max(A.kind_of(sx), A.kind_of(sy))
    nx, ny = A.numel(sx), A.numel(sy)
    if nx != ny:
        raise ValueError("slope arrays have different lengths")
    if kind == A.KIND_CUDA:
        import torch
        rays = torch.stack([torch.full_like(sx, float(x0), dtype=torch.float64),
                            torch.full_like(sx, float(y0), dtype=torch.float64),
                            sx.to(torch.float64), sy.to(torch.float64),
                            torch.ones_like(sx, dtype=torch.float64)], dim=-1)
    else:
        sxn, syn = A.to_host_f64(sx), A.to_host_f64(sy)
        rays = np.stack([np.full(nx, float(x0)), np.full(nx, float(y0)), sxn, syn, np.ones(nx)], axis=-1)
        if nx == 0:
            return np.zeros((4, 0))
    out = _apply(rays, _np(transfer_matrix)[None])
    res = out[:, 0, :4].T
    if kind == A.KIND_TORCH_CPU:
        import torch
        return torch.as_tensor(res)
    return res
