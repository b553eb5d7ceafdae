"""``KrivanekCoeffs`` (reference ``src/temgym_core/aberrations.py:5-31``).

The aberration function itself (``W_krivanek`` / ``grad_W_krivanek``,
aberrations.py:42-108) is evaluated inside the CUDA ray kernel
(``csrc/trace.cu``); the 25 coefficients travel in the model descriptor in
the field order below.  ``W_krivanek`` / ``grad_W_krivanek`` called directly run the same device
code on arrays of slopes (``tg_krivanek_f64``).
"""
from dataclasses import dataclass, fields

import numpy as np


@dataclass(unsafe_hash=True)   # hashable: components holding it are dictionary keys of run_with_grads
class KrivanekCoeffs:
    C10: float = 0.0
    C12: float = 0.0
    phi12: float = 0.0
    C21: float = 0.0
    phi21: float = 0.0
    C23: float = 0.0
    phi23: float = 0.0
    C30: float = 0.0
    C32: float = 0.0
    phi32: float = 0.0
    C34: float = 0.0
    phi34: float = 0.0
    C41: float = 0.0
    phi41: float = 0.0
    C43: float = 0.0
    phi43: float = 0.0
    C45: float = 0.0
    phi45: float = 0.0
    C50: float = 0.0
    C52: float = 0.0
    phi52: float = 0.0
    C54: float = 0.0
    phi54: float = 0.0
    C56: float = 0.0
    phi56: float = 0.0

    def as_tuple(self):
        return tuple(float(getattr(self, f.name)) for f in fields(self))


# harmonic terms (m, phi0 field) in the kernel's order (include/temgym_b200.h, tg_krivanek_f64)
_TERMS = ((2, "phi12"), (1, "phi21"), (3, "phi23"), (2, "phi32"), (4, "phi34"), (1, "phi41"),
          (3, "phi43"), (5, "phi45"), (2, "phi52"), (4, "phi54"), (6, "phi56"))


def krivanek_param_block(p) -> tuple:
    """25 coefficients + 11 (cos, sin)(m phi0) pairs: the parameter block the kernels read."""
    if isinstance(p, dict):
        p = KrivanekCoeffs(**p)
    trig = []
    for m, name in _TERMS:
        ph0 = float(getattr(p, name))
        trig += [float(np.cos(m * ph0)), float(np.sin(m * ph0))]
    return tuple(float(v) for v in p.as_tuple()) + tuple(trig)


def _krivanek_call(alpha_x, alpha_y, p, want):
    import torch
    from . import _arrays as A
    from . import _lib as L
    lib = L.load()
    kind = max(A.kind_of(alpha_x), A.kind_of(alpha_y))
    shape = A.shape_of(alpha_x) if A.kind_of(alpha_x) != A.KIND_SCALAR else A.shape_of(alpha_y)
    dev = A.cuda_device_of((alpha_x, alpha_y)) or torch.device("cuda", A.current_device_index())
    n = max(A.numel(alpha_x), A.numel(alpha_y))

    def dev_arr(v):
        t = A.to_device_f64(v, dev)
        return t.expand(n).contiguous() if t.numel() == 1 and n > 1 else t
    ax, ay = dev_arr(alpha_x), dev_arr(alpha_y)
    outs = [torch.empty(n, dtype=torch.float64, device=dev) if w else None for w in want]
    with torch.cuda.device(dev):
        L.check(lib.tg_krivanek_f64(n, ax.data_ptr(), ay.data_ptr(), L.dbl_array(krivanek_param_block(p)),
                                    *[o.data_ptr() if o is not None else None for o in outs],
                                    A.current_stream_ptr(dev)), "tg_krivanek_f64")

    def fin(t):
        t = t.reshape(shape)
        if kind == A.KIND_CUDA:
            return t
        if kind == A.KIND_SCALAR:
            return float(t.reshape(-1)[0].item())
        return t.cpu() if kind == A.KIND_TORCH_CPU else t.cpu().numpy()
    return [fin(o) for o in outs if o is not None]


def grad_W_krivanek(alpha_x, alpha_y, p):
    """``(dW/d alpha_x, dW/d alpha_y)`` of the Krivanek aberration function (aberrations.py:63-108)."""
    dWx, dWy = _krivanek_call(alpha_x, alpha_y, p, (False, True, True))
    return dWx, dWy


def W_krivanek(alpha, phi, p):
    """The aberration function W(alpha, phi) in polar slope coordinates (aberrations.py:51-60)."""
    if hasattr(alpha, "detach") or hasattr(phi, "detach"):
        import torch
        alpha, phi = torch.as_tensor(alpha), torch.as_tensor(phi)
        ax, ay = alpha * torch.cos(phi), alpha * torch.sin(phi)
    else:
        ax, ay = np.asarray(alpha) * np.cos(phi), np.asarray(alpha) * np.sin(phi)
    return _krivanek_call(ax, ay, p, (True, False, False))[0]
