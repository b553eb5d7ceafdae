"""Array plumbing between the Python call surface and the C ABI.

The reference hands ``jax.numpy`` arrays (or Python floats) to its functions.  Here the
same calls accept Python scalars, numpy arrays, and torch tensors (CPU or CUDA) and
answer in kind:

* any CUDA tensor among the inputs  -> device path (``tg_*`` with device pointers on
  torch's current stream), results are CUDA tensors;
* otherwise                         -> host path (``tg_*_host``: the library does the
  H2D copy, the kernels and the D2H copy), results are numpy arrays / CPU tensors /
  Python floats.

torch is used for device memory and streams only.
"""
from __future__ import annotations

import numpy as np

try:  # torch is the device-memory allocator; the host path works without it
    import torch
except Exception:  # pragma: no cover
    torch = None

KIND_SCALAR, KIND_NUMPY, KIND_TORCH_CPU, KIND_CUDA = 0, 1, 2, 3


def kind_of(v) -> int:
    if torch is not None and isinstance(v, torch.Tensor):
        return KIND_CUDA if v.is_cuda else KIND_TORCH_CPU
    if isinstance(v, np.ndarray) and v.ndim > 0:
        return KIND_NUMPY
    if isinstance(v, (list, tuple)):
        return KIND_NUMPY
    return KIND_SCALAR


def numel(v) -> int:
    k = kind_of(v)
    if k == KIND_SCALAR:
        return 1
    if k in (KIND_TORCH_CPU, KIND_CUDA):
        return int(v.numel())
    return int(np.asarray(v).size)


def shape_of(v):
    k = kind_of(v)
    if k == KIND_SCALAR:
        return ()
    if k in (KIND_TORCH_CPU, KIND_CUDA):
        return tuple(v.shape)
    return tuple(np.asarray(v).shape)


def to_float(v) -> float:
    if torch is not None and isinstance(v, torch.Tensor):
        return float(v.item())
    return float(np.asarray(v).reshape(-1)[0]) if isinstance(v, np.ndarray) else float(v)


def to_host_f64(v) -> np.ndarray:
    """Flat contiguous float64 numpy array (no copy when already so)."""
    if type(v) is np.ndarray and v.dtype == np.float64 and v.flags.c_contiguous:
        return v if v.ndim == 1 else v.reshape(-1)
    if torch is not None and isinstance(v, torch.Tensor):
        v = v.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(-1))


def to_device_f64(v, device):
    """Flat contiguous float64 CUDA tensor on ``device``."""
    if isinstance(v, torch.Tensor):
        t = v.detach().to(device=device, dtype=torch.float64)
    else:
        t = torch.as_tensor(np.asarray(v, dtype=np.float64), device=device)
    return t.reshape(-1).contiguous()


def cuda_device_of(values):
    for v in values:
        if kind_of(v) == KIND_CUDA:
            return v.device
    return None


def current_device_index() -> int:
    if torch is not None and torch.cuda.is_available():
        return int(torch.cuda.current_device())
    return 0


def current_stream_ptr(device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


def from_host(arr: np.ndarray, kind: int, shape=None):
    """Return a host result in the caller's array kind."""
    if shape is not None:
        arr = arr.reshape(shape)
    if kind == KIND_TORCH_CPU:
        return torch.from_numpy(np.ascontiguousarray(arr))
    return arr
