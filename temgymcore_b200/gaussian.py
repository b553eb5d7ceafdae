"""Gaussian-beamlet wave optics (reference ``src/temgym_core/gaussian.py``).

``make_gaussian_image`` / ``evaluate_gaussian_input_image`` and the two flat-array
entry points ``propagate_misaligned_gaussian_jax_scan`` /
``evaluate_misaligned_input_gaussian_jax_scan`` keep the reference's names, argument
order and shapes; the work runs in CUDA kernels:

* central rays + per-ray 5x5 ABCD          -> ``tg_trace_f64``            (csrc/trace.cu)
* ``Q_inv``, ``k``, ``phase_offset``       -> ``tg_gaussian_qinv_f64`` ... (csrc/coeffs.cu)
* per-beamlet complex quadratic            -> ``tg_beamlet_coeffs_*``      (csrc/coeffs.cu)
* sum over beamlets on every pixel         -> ``tg_field_sum_grid``        (csrc/field.cu)

``batch_size`` is accepted for signature compatibility; the chunking it controlled
(``map_reduce``, gaussian.py:340-369) is replaced by on-chip tiling.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, fields
from typing import Any

import numpy as np

from . import _arrays as A
from . import _lib as L
from .grid import Grid
from .ray import RAY_FIELDS, Ray
from .run import _check_propagator, compile_model as _compile_any, require_scalar_params  # noqa: F401


def compile_model(model):
    return require_scalar_params(_compile_any(model), "make_gaussian_image")

DEFAULT_CULL_BITS = 40
"""Beamlets whose envelope over a whole pixel tile is below 2**-40 of the brightest
beamlet peak on the detector are skipped for that tile (far below the fp32 evaluation
noise of ~2**-20).  Pass ``cull_bits=0`` for the dense sum."""


# closed forms used as analytic anchors by the reference tests (gaussian.py:13-32, 99-110)
def w_z(w0, z, z_r):
    return w0 * np.sqrt(1 + (z / z_r) ** 2)


def zR(w0, wavelength):
    return (np.pi * w0 ** 2) / wavelength


def R(z, z_r):
    if abs(z) < 1e-10:
        return np.inf
    return z * (1 + (z_r / z) ** 2)


def q_inv(z, w0, wl):
    z_r = zR(w0, wl)
    if abs(z) < 1e-10:
        return 1j * wl / (np.pi * w0 ** 2)
    return -1.0 / R(z, z_r) + 1j * wl / (np.pi * w_z(w0, z, z_r) ** 2)


def gaussian_beam(x, y, q_inv, k, offset_x=0, offset_y=0):
    return np.exp(1j * k * ((x + offset_x) ** 2 + (y + offset_y) ** 2) / 2 * q_inv)


def decompose_Q_inv(Q_inv, wavelength, eps=1e-12):
    """``(waist_x, waist_y, r_x, r_y, theta)`` from 2x2 complex ``Q_inv`` (batched)
    (gaussian.py:35-89): principal axes from the symmetric imaginary part (``eigh``), right-handed
    eigenvectors, larger waist first, waists from Im, radii from Re of the rotated diagonal.
    A CUDA tensor is decomposed on its device (``tg_decompose_qinv_f64``, closed-form 2x2 eigenvectors) and the
    results are CUDA tensors; anything else takes the host numpy route of the reference."""
    if A.kind_of(Q_inv) == A.KIND_CUDA:
        import torch
        lib = L.load()
        dev = Q_inv.device
        Qc = Q_inv.to(torch.complex128).contiguous()
        lead = tuple(Qc.shape[:-2])
        n = int(np.prod(lead)) if lead else 1
        wl = A.to_device_f64(wavelength, dev)
        if wl.numel() not in (1, n):
            raise ValueError("wavelength must be a scalar or match the batch of Q_inv")
        outs = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(5)]
        with torch.cuda.device(dev):
            L.check(lib.tg_decompose_qinv_f64(n, torch.view_as_real(Qc).data_ptr(), wl.data_ptr(),
                                              int(wl.numel() == 1), float(eps), *[o.data_ptr() for o in outs],
                                              A.current_stream_ptr(dev)), "tg_decompose_qinv_f64")
        return tuple(o.reshape(lead) for o in outs)
    Q = np.asarray(_to_np(Q_inv), dtype=np.complex128)
    Sm = np.imag(Q)
    Sm = 0.5 * (Sm + np.swapaxes(Sm, -1, -2))
    _, ev = np.linalg.eigh(Sm)

    def right_handed(e):
        sgn = np.where(np.linalg.det(e) < 0, -1.0, 1.0)
        e = e.copy()
        e[..., :, 1] *= sgn[..., None]
        return e

    ev = right_handed(ev)
    Qd = np.swapaxes(ev, -1, -2) @ Q @ ev
    qd = np.stack([Qd[..., 0, 0], Qd[..., 1, 1]], axis=-1)

    def waists_of(q):
        im = np.imag(q)
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.sqrt(np.where(np.abs(im) > eps, np.abs(wavelength / (np.pi * im)), np.inf))

    swap = waists_of(qd)[..., 0] < waists_of(qd)[..., 1]
    qd = np.where(swap[..., None], qd[..., ::-1], qd)
    ev = right_handed(np.where(swap[..., None, None], ev[..., :, ::-1], ev))
    w = waists_of(qd)
    re = np.real(qd)
    with np.errstate(divide="ignore", invalid="ignore"):
        radii = np.where(np.abs(re) > eps, 1.0 / re, np.inf)
    theta = np.arctan2(ev[..., 1, 0], ev[..., 0, 0])
    return w[..., 0], w[..., 1], radii[..., 0], radii[..., 1], theta


@dataclass(frozen=True, kw_only=True, eq=False)
class GaussianRay(Ray):
    """Ray + Gaussian beam parameters (gaussian.py:113-177)."""
    amplitude: Any
    waist_xy: Any
    radii_of_curv: Any
    wavelength: Any
    theta: Any

    def derive(self, **updates):
        import dataclasses

        def resolve(v):
            return v(self) if callable(v) else v
        return dataclasses.replace(self, **{k: resolve(v) for k, v in updates.items()})

    def to_ray(self):
        return Ray(x=self.x, y=self.y, dx=self.dx, dy=self.dy, z=self.z,
                   pathlength=self.pathlength, _one=self._one)

    def to_vector(self):
        def v1(v):
            if A.kind_of(v) in (A.KIND_TORCH_CPU, A.KIND_CUDA):
                return v.reshape(-1) if v.ndim == 0 else v
            return np.atleast_1d(np.asarray(v, dtype=np.float64))
        return type(self)(**{f.name: v1(getattr(self, f.name)) for f in fields(self)})

    @property
    def q_inv(self):
        """``(1/q_x, 1/q_y)`` on the principal axes: ``-1/R + i lambda / (pi w^2)``, ``R = inf`` giving a
        purely imaginary value (gaussian.py:138-155).  O(n) input preparation on the arrays' own side
        (numpy or torch); the rotated 2x2 ``Q_inv`` below comes from the CUDA kernel."""
        w, Rc, wl = self.waist_xy, self.radii_of_curv, self.wavelength
        if hasattr(w, "detach") or hasattr(Rc, "detach") or hasattr(wl, "detach"):
            import torch
            w, Rc = torch.as_tensor(w, dtype=torch.float64), torch.as_tensor(Rc, dtype=torch.float64)
            wl = torch.as_tensor(wl, dtype=torch.float64, device=w.device)
            im = wl[..., None] / (torch.pi * w ** 2) if wl.ndim else wl / (torch.pi * w ** 2)
            re = torch.where(torch.isinf(Rc), torch.zeros_like(Rc), -1.0 / Rc)
            q = torch.complex(re, im)
        else:
            w, Rc, wl = np.asarray(w, float), np.asarray(Rc, float), np.asarray(wl, float)
            wl = wl[..., None] if wl.ndim else wl
            with np.errstate(divide="ignore"):
                re = np.where(np.isinf(Rc), 0.0, -1.0 / Rc)
            q = re + 1j * wl / (np.pi * w ** 2)          # the reference's expression, term for term
        return q[..., 0], q[..., 1]

    @property
    def Q_inv(self):
        """(n, 2, 2) complex ``R diag(1/q_x, 1/q_y) R^T`` (gaussian.py:138-177), computed by
        ``tg_gaussian_qinv_f64``; a CUDA complex128 tensor."""
        import torch
        dev = _device_for(self)
        g = _beamlet_arrays(self, dev)
        return torch.view_as_complex(_qinv(g, dev).reshape(-1, 2, 2, 2))


# ------------------------------------------------------------------------------ helpers
def _device_for(obj):
    import torch
    vals = [getattr(obj, f.name) for f in fields(obj)] if hasattr(obj, "__dataclass_fields__") else list(obj)
    dev = A.cuda_device_of(vals)
    return dev if dev is not None else torch.device("cuda", A.current_device_index())


def _beamlet_arrays(g: GaussianRay, dev):
    """All GaussianRay leaves as flat fp64 CUDA tensors of one common length nb."""
    import torch
    n = 1
    for f in RAY_FIELDS + ("amplitude", "wavelength", "theta"):
        n = max(n, A.numel(getattr(g, f)))
    n = max(n, A.numel(g.waist_xy) // 2, A.numel(g.radii_of_curv) // 2)

    def vec(v):
        t = A.to_device_f64(v, dev)
        return t.expand(n).contiguous() if t.numel() == 1 and n > 1 else t

    def vec2(v):
        t = A.to_device_f64(v, dev).reshape(-1, 2)
        return (t.expand(n, 2) if t.shape[0] == 1 and n > 1 else t).contiguous()

    d = {f: vec(getattr(g, f)) for f in RAY_FIELDS + ("amplitude", "wavelength", "theta")}
    d["waist_xy"] = vec2(g.waist_xy)
    d["radii_of_curv"] = vec2(g.radii_of_curv)
    for k, t in d.items():
        if t.shape[0] != n:
            raise ValueError(f"GaussianRay field {k!r} has {t.shape[0]} entries, expected {n}")
    d["n"] = n
    return d


def _qinv(g, dev):
    import torch
    lib = L.load()
    n = g["n"]
    q = torch.empty((n, 8), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.tg_gaussian_qinv_f64(n, g["waist_xy"].data_ptr(), g["radii_of_curv"].data_ptr(),
                                         g["wavelength"].data_ptr(), g["theta"].data_ptr(),
                                         q.data_ptr(), A.current_stream_ptr(dev)),
                "tg_gaussian_qinv_f64")
    return q


def _wave_numbers(g, dev):
    import torch
    lib = L.load()
    n = g["n"]
    k = torch.empty(n, dtype=torch.float64, device=dev)
    p0 = torch.empty(n, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.tg_wave_numbers_f64(n, g["wavelength"].data_ptr(), g["pathlength"].data_ptr(),
                                        k.data_ptr(), p0.data_ptr(), A.current_stream_ptr(dev)),
                "tg_wave_numbers_f64")
    return k, p0


def _field_sum_grid(poly, nb, grid: Grid, dev, row0=0, nrows=None, out_dtype=None, cull_bits=None,
                    count_evals=False, method="sfu"):
    """Sum the beamlet field on rows [row0, row0+nrows) of ``grid``.

    method: "sfu" = tiled SFU kernel (any beamlets); "tensor" = tcgen05 GEMM for separable
    beamlets (raises if they are not); "auto" = tensor when separable, else sfu."""
    import torch
    lib = L.load()
    if method not in L.TG_METHOD:
        raise ValueError(f"method must be one of {sorted(L.TG_METHOD)}")
    H, W = int(grid.shape[0]), int(grid.shape[1])
    nrows = H - row0 if nrows is None else nrows
    out_dtype = torch.complex128 if out_dtype is None else out_dtype
    if out_dtype not in (torch.complex64, torch.complex128):
        raise ValueError("out_dtype must be torch.complex64 or torch.complex128")
    cull = DEFAULT_CULL_BITS if cull_bits is None else int(cull_bits)
    out = torch.empty((nrows, W), dtype=out_dtype, device=dev)
    nev = C.c_longlong(0)
    with torch.cuda.device(dev):
        if count_evals:
            if method == "tensor":
                raise ValueError("count_evals is a property of the SFU kernel")
            L.check(lib.tg_field_sum_grid(nb, poly.data_ptr() if nb else None,
                                          L.dbl_array(grid.px2m_affine), H, W, row0, nrows,
                                          out.data_ptr(), int(out_dtype == torch.complex128), cull,
                                          C.byref(nev), A.current_stream_ptr(dev)), "tg_field_sum_grid")
        else:
            L.check(lib.tg_field_sum(nb, poly.data_ptr() if nb else None, L.dbl_array(grid.px2m_affine),
                                     H, W, row0, nrows, out.data_ptr(),
                                     int(out_dtype == torch.complex128), cull, L.TG_METHOD[method],
                                     A.current_stream_ptr(dev)), "tg_field_sum")
    return (out, int(nev.value)) if count_evals else out


def _result_kind(obj) -> int:
    vals = [getattr(obj, f.name) for f in fields(obj)]
    return max(A.kind_of(v) for v in vals)


def _pinned_empty(shape, dtype):
    """Pinned host buffer from torch's caching host allocator (D2H at full PCIe rate)."""
    import torch
    return torch.empty(shape, dtype=dtype, pin_memory=True)


def _finish(out, kind):
    if kind == A.KIND_CUDA:
        return out
    import torch
    host = _pinned_empty(tuple(out.shape), out.dtype)
    host.copy_(out, non_blocking=True)
    torch.cuda.current_stream(out.device).synchronize()
    return host if kind == A.KIND_TORCH_CPU else host.numpy()


_PACK_ORDER = RAY_FIELDS + ("amplitude", "waist_xy", "radii_of_curv", "wavelength", "theta")


def pack_beamlets_pinned(gaussian_rays):
    """The same ``GaussianRay`` with every field a view into ONE page-locked host slab, in the order the
    host-buffer call uploads them (rays, amplitude, waist_xy, radii_of_curv, wavelength, theta): the H2D of
    ``make_gaussian_image`` / ``tg_make_gaussian_image_host`` is then a single PCIe copy of 112 bytes per
    beamlet instead of twelve small ones.  Fields are numpy views; writing into them updates the slab."""
    import torch
    g = gaussian_rays
    n = 1
    for f in RAY_FIELDS + ("amplitude", "wavelength", "theta"):
        n = max(n, A.numel(getattr(g, f)))
    n = max(n, A.numel(g.waist_xy) // 2, A.numel(g.radii_of_curv) // 2)
    slab = torch.empty(14 * n, dtype=torch.float64, pin_memory=True).numpy()
    views, off = {}, 0
    for f in _PACK_ORDER:
        width = 2 if f in ("waist_xy", "radii_of_curv") else 1
        h = A.to_host_f64(getattr(g, f))
        if h.size == width and n > 1:
            h = np.broadcast_to(h.reshape(1, width), (n, width)).reshape(-1)
        if h.size != n * width:
            raise ValueError(f"GaussianRay field {f!r} has {h.size} entries, expected {n * width}")
        v = slab[off:off + n * width]
        v[:] = h
        views[f] = v.reshape(n, 2) if width == 2 else v
        off += n * width
    packed = type(g)(**views)
    # the host-buffer call's pointer arguments, resolved once (twelve array-interface lookups are ~25 us of
    # Python per call otherwise); `dataclasses.replace` / `derive` make a new object without this memo
    hp = _host_ptr
    object.__setattr__(packed, "_tg_host_args", (
        n, tuple(hp(views[f]) for f in RAY_FIELDS),
        tuple(hp(views[f]) for f in ("amplitude", "waist_xy", "radii_of_curv", "wavelength", "theta")), slab))
    return packed


def _host_ptr(a: np.ndarray) -> int:
    return a.__array_interface__["data"][0]


def make_gaussian_image_host(gaussian_rays, model, *, cull_bits=None, out_dtype=None, row0=0,
                             nrows=None, device=None, method="auto"):
    """``make_gaussian_image`` for HOST inputs through the single host-buffer C-ABI call
    ``tg_make_gaussian_image_host`` (H2D copies, all kernels, D2H copy inside the call)."""
    import torch
    lib = L.load()
    g = gaussian_rays
    grid = model[-1]
    H, W = int(grid.shape[0]), int(grid.shape[1])
    nrows = H - row0 if nrows is None else nrows
    memo = getattr(g, "_tg_host_args", None)       # pack_beamlets_pinned resolved the pointers already
    if memo is not None:
        n, ray_ptrs, (p_amp, p_waist, p_radii, p_wl, p_th), _slab = memo
    else:
        n = 1
        for f in RAY_FIELDS + ("amplitude", "wavelength", "theta"):
            n = max(n, A.numel(getattr(g, f)))
        n = max(n, A.numel(g.waist_xy) // 2, A.numel(g.radii_of_curv) // 2)

        def vec(v, width=1):
            h = A.to_host_f64(v)
            if h.size == width and n > 1:
                h = np.ascontiguousarray(np.broadcast_to(h.reshape(1, width), (n, width)).reshape(-1))
            if h.size != n * width:
                raise ValueError("GaussianRay fields have mismatched sizes")
            return h

        keep = [vec(getattr(g, f)) for f in RAY_FIELDS]
        keep += [vec(g.amplitude), vec(g.waist_xy, 2), vec(g.radii_of_curv, 2), vec(g.wavelength), vec(g.theta)]
        hp = _host_ptr
        ray_ptrs = tuple(hp(r) for r in keep[:7])
        p_amp, p_waist, p_radii, p_wl, p_th = (hp(r) for r in keep[7:])
    out_dtype = torch.complex128 if out_dtype is None else out_dtype
    out = _pinned_empty((nrows, W), out_dtype)
    cull = DEFAULT_CULL_BITS if cull_bits is None else int(cull_bits)
    cm = compile_model(model)
    dev = A.current_device_index() if device is None else int(device)
    L.check(lib.tg_make_gaussian_image_host(
        C.byref(cm), n, L.ptr_array(ray_ptrs), p_amp, p_waist, p_radii, p_wl, p_th,
        L.dbl_array(grid.px2m_affine), H, W, row0, nrows, out.data_ptr(),
        int(out_dtype == torch.complex128), cull, L.TG_METHOD[method], dev),
        "tg_make_gaussian_image_host")
    return out


def beamlet_polynomials(gaussian_rays: GaussianRay, model):
    """Trace the central rays (ABCD per ray) and build the six complex coefficients per
    beamlet: the device-side front half of ``make_gaussian_image`` (gaussian.py:227-255).
    Returns ``(poly (nb,12) CUDA tensor, nb, device)``."""
    import torch
    lib = L.load()
    dev = _device_for(gaussian_rays)
    g = _beamlet_arrays(gaussian_rays, dev)
    n = g["n"]
    cm = compile_model(model)
    abcd = torch.empty((n, 25), dtype=torch.float64, device=dev)
    rin = L.tg_ray_in()
    for i, f in enumerate(RAY_FIELDS):
        rin.ptr[i] = g[f].data_ptr()
    poly = torch.empty((n, 12), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        st = A.current_stream_ptr(dev)
        L.check(lib.tg_trace_f64(C.byref(cm), n, C.byref(rin), L.ptr_array([None] * 7),
                                 abcd.data_ptr(), L.TG_JAC_ABCD5, st), "tg_trace_f64")
        q = _qinv(g, dev)
        k, p0 = _wave_numbers(g, dev)
        L.check(lib.tg_beamlet_coeffs_abcd_f64(n, g["amplitude"].data_ptr(), p0.data_ptr(),
                                               q.data_ptr(), abcd.data_ptr(), g["x"].data_ptr(),
                                               g["y"].data_ptr(), g["dx"].data_ptr(),
                                               g["dy"].data_ptr(), k.data_ptr(), poly.data_ptr(), st),
                "tg_beamlet_coeffs_abcd_f64")
    return poly, n, dev


# ------------------------------------------------------------------------------ public API
# small einsum helpers with the reference's index conventions (per beamlet: the reference vmaps them)
def _einsum(spec, *ops):
    if any(hasattr(o, "detach") for o in ops):
        import torch
        return torch.einsum(spec, *ops)
    return np.einsum(spec, *ops)


def matrix_vector_mul(M, v):                  # gaussian.py:180-187
    return _einsum("ij,j->i", M, v)


def matrix_matrix_mul(M1, M2):                # gaussian.py:190-197
    return _einsum("ij,jk->ik", M1, M2)


def matrix_quadratic_mul(v, M):               # gaussian.py:200-207
    return _einsum("i,ij,j->", v, M, v)


def matrix_linear_mul(v, M, w):               # gaussian.py:210-218
    return _einsum("i,ij,nj->n", v, M, w)


def matrix_matrix_matrix_mul(M1, M2, M3):     # gaussian.py:221-222
    return _einsum("nij,njk,npk->nip", M1, M2, M3)


def Qinv_ABCD(Qinv, A_, B, C_, D):
    """``solve(A + B Qinv, C + D Qinv)`` (gaussian.py:92-96) -- tiny host helper (numpy);
    inside the field sum this is evaluated per beamlet by the coefficient kernel."""
    Qinv = np.asarray(_to_np(Qinv))
    lhs = np.asarray(_to_np(A_)) + np.asarray(_to_np(B)) @ Qinv
    rhs = np.asarray(_to_np(C_)) + np.asarray(_to_np(D)) @ Qinv
    return np.linalg.solve(lhs, rhs)


def _to_np(v):
    return v.detach().cpu().numpy() if hasattr(v, "detach") else v


def make_gaussian_image(gaussian_rays, model, batch_size=128, *, cull_bits=None, out_dtype=None,
                        method="auto"):
    """Field of all beamlets on the detector ``model[-1]`` -> ``(H, W)`` complex128
    (gaussian.py:225-273).  ``method``: "auto" (tensor cores when the beamlets are separable
    on this grid, else the SFU kernel), "sfu", or "tensor"."""
    rays = gaussian_rays
    assert isinstance(rays, GaussianRay)
    grid = model[-1]
    assert isinstance(grid, Grid)
    kind = _result_kind(rays)
    if kind != A.KIND_CUDA:
        out = make_gaussian_image_host(rays, model, cull_bits=cull_bits, out_dtype=out_dtype,
                                       method=method)
        return out if kind == A.KIND_TORCH_CPU else out.numpy()
    return make_gaussian_image_device(rays, model, cull_bits=cull_bits, out_dtype=out_dtype, method=method)


def make_gaussian_image_device(gaussian_rays, model, *, cull_bits=None, out_dtype=None, method="auto",
                               row0=0, nrows=None):
    """``make_gaussian_image`` for device-resident inputs through the single C-ABI call
    ``tg_make_gaussian_image_f64`` (everything enqueued on torch's current stream, no host
    synchronisation); returns a CUDA tensor of rows ``[row0, row0+nrows)``."""
    import torch
    lib = L.load()
    grid = model[-1]
    H, W = int(grid.shape[0]), int(grid.shape[1])
    nrows = H - row0 if nrows is None else nrows
    dev = _device_for(gaussian_rays)
    g = _beamlet_arrays(gaussian_rays, dev)
    out_dtype = torch.complex128 if out_dtype is None else out_dtype
    if out_dtype not in (torch.complex64, torch.complex128):
        raise ValueError("out_dtype must be torch.complex64 or torch.complex128")
    cull = DEFAULT_CULL_BITS if cull_bits is None else int(cull_bits)
    out = torch.empty((nrows, W), dtype=out_dtype, device=dev)
    cm = compile_model(model)
    with torch.cuda.device(dev):
        L.check(lib.tg_make_gaussian_image_f64(
            C.byref(cm), g["n"], L.ptr_array([g[f].data_ptr() for f in RAY_FIELDS]),
            g["amplitude"].data_ptr(), g["waist_xy"].data_ptr(), g["radii_of_curv"].data_ptr(),
            g["wavelength"].data_ptr(), g["theta"].data_ptr(), L.dbl_array(grid.px2m_affine), H, W,
            row0, nrows, out.data_ptr(), int(out_dtype == torch.complex128), cull, L.TG_METHOD[method],
            A.current_stream_ptr(dev)), "tg_make_gaussian_image_f64")
    return out


def auto_dispatch(gaussian_rays, model, cull_bits=None) -> str:
    """What ``method="auto"`` decides for these beamlets on ``model[-1]``: ``"tensor"`` (dense GEMM), ``"tensor_binned"``
    (separable and sparse: tile-binned GEMM) or ``"sfu"``
    (``tg_field_sum_verdict``: the device-side separability + cost verdict, read back once; synchronises)."""
    lib = L.load()
    grid = model[-1]
    poly, nb, dev = beamlet_polynomials(gaussian_rays, model)
    cull = DEFAULT_CULL_BITS if cull_bits is None else int(cull_bits)
    use = C.c_int(0)
    import torch
    with torch.cuda.device(dev):
        L.check(lib.tg_field_sum_verdict(nb, poly.data_ptr() if nb else None, L.dbl_array(grid.px2m_affine),
                                         int(grid.shape[0]), int(grid.shape[1]), cull, C.byref(use),
                                         A.current_stream_ptr(dev)), "tg_field_sum_verdict")
    return {0: "sfu", 1: "tensor", 2: "tensor_binned"}[use.value]


class GaussianImagePlan:
    """``make_gaussian_image`` captured once into a CUDA graph and replayed.

    For repeated imaging with fixed shapes (parameter sweeps, scan positions, optimisation loops):
    the ~12 small launches of one image (ray kernel, Q_inv, coefficients, factors, GEMM ...) are
    replayed as ONE graph launch, which removes the host launch gaps that dominate once the field
    sum itself takes < 0.5 ms.  Inputs live in the plan's static CUDA buffers: ``update(rays)``
    copies new beamlet parameters in, ``run()`` replays the graph and returns the (static) output
    tensor.  The model (component parameters) is baked into the graph; build a new plan to change it.
    """

    def __init__(self, gaussian_rays, model, *, cull_bits=None, out_dtype=None, method="auto",
                 freeze_dispatch: bool = True):
        """``freeze_dispatch`` (default): ``method="auto"`` is resolved ONCE, here, from the plan's first beamlets
        (``auto_dispatch``) and the graph holds the launches of that path only; ``self.method`` tells which.  The
        tensor path keeps its device-side separability guard: if ``update()`` later brings beamlets that are not
        separable on this grid the image is filled with NaN -- build a new plan (or pass ``freeze_dispatch=False``
        to keep both paths and the device-side verdict in the graph) when the beamlet class can change."""
        import torch
        self._model = list(model)
        if method == "auto" and freeze_dispatch:
            method = auto_dispatch(gaussian_rays, self._model, cull_bits)
        self.method = method
        self._kw = dict(cull_bits=cull_bits, out_dtype=out_dtype, method=method)
        dev = _device_for(gaussian_rays)
        self.device = dev
        # every field (Python scalars too) becomes an expanded static CUDA tensor BEFORE the capture: a scalar
        # left in place would turn into a pageable H2D copy inside the graph
        garr = _beamlet_arrays(gaussian_rays, dev)
        self._static = type(gaussian_rays)(**{f.name: garr[f.name].clone() for f in fields(gaussian_rays)})
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                      # warm-up outside the capture
            for _ in range(2):
                make_gaussian_image_device(self._static, self._model, **self._kw)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._out = make_gaussian_image_device(self._static, self._model, **self._kw)

    def update(self, gaussian_rays):
        """Copy new beamlet parameters (same shapes) into the plan's static buffers."""
        for f in fields(gaussian_rays):
            dst, src = getattr(self._static, f.name), getattr(gaussian_rays, f.name)
            t = A.to_device_f64(src, self.device)
            if t.numel() != dst.numel():
                if t.numel() not in (1, 2) or dst.numel() % t.numel():
                    raise ValueError(f"GaussianImagePlan.update: field {f.name!r} has {t.numel()} entries, "
                                     f"the plan was built for {dst.numel()}")
                t = t.reshape(1, -1).expand(dst.numel() // t.numel(), t.numel())
            dst.copy_(t.reshape(dst.shape), non_blocking=True)
        return self

    def run(self):
        self._graph.replay()
        return self._out

    __call__ = run


def replace_fields(obj, fn):
    import dataclasses
    return dataclasses.replace(obj, **{f.name: fn(getattr(obj, f.name)) for f in fields(obj)})


def evaluate_gaussian_input_image(gaussian_rays, grid, batch_size=128, *, cull_bits=None,
                                  out_dtype=None, method="auto"):
    """Input-plane field of all beamlets on ``grid`` (gaussian.py:372-399)."""
    import torch
    lib = L.load()
    rays = gaussian_rays
    assert isinstance(rays, GaussianRay)
    kind = _result_kind(rays)
    dev = _device_for(rays)
    g = _beamlet_arrays(rays, dev)
    n = g["n"]
    q = _qinv(g, dev)
    k, p0 = _wave_numbers(g, dev)
    r1m = torch.stack([g["x"], g["y"]], dim=-1).contiguous()
    th = torch.stack([g["dx"], g["dy"]], dim=-1).contiguous()
    poly = torch.empty((n, 12), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.tg_input_coeffs_f64(n, g["amplitude"].data_ptr(), p0.data_ptr(), q.data_ptr(),
                                        r1m.data_ptr(), th.data_ptr(), k.data_ptr(), poly.data_ptr(),
                                        A.current_stream_ptr(dev)), "tg_input_coeffs_f64")
    out = _field_sum_grid(poly, n, grid, dev, out_dtype=out_dtype, cull_bits=cull_bits, method=method)
    return _finish(out, kind)


def _flat_inputs(dev, *arrs):
    import torch
    out = []
    for a in arrs:
        if isinstance(a, torch.Tensor):
            t = a.detach().to(dev)
        else:
            t = torch.as_tensor(np.asarray(a), device=dev)
        if t.is_complex():
            t = torch.view_as_real(t.to(torch.complex128).contiguous())
        out.append(t.to(torch.float64).contiguous())
    return out


def _points_or_grid(poly, nb, r, dev, grid, cull_bits, kind):
    import torch
    lib = L.load()
    if grid is not None:
        out = _field_sum_grid(poly, nb, grid, dev, cull_bits=cull_bits).reshape(-1)
        return _finish(out, kind)
    (rt,) = _flat_inputs(dev, r)
    rt = rt.reshape(-1, 2).contiguous()
    npts = rt.shape[0]
    out = torch.empty(npts, dtype=torch.complex128, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.tg_field_sum_points(nb, poly.data_ptr() if nb else None, npts, rt.data_ptr(),
                                        out.data_ptr(), 1, A.current_stream_ptr(dev)),
                "tg_field_sum_points")
    return _finish(out, kind)


def propagate_misaligned_gaussian_jax_scan(amp, phase_offset, Q1_inv, A_, B, C_, D, e, f, r1m,
                                           theta1m, k, r2, batch_size=128, *, grid=None,
                                           cull_bits=None):
    """Flat-array field sum at the observation points ``r2 (npix, 2)`` -> ``(npix,)``
    complex128 (gaussian.py:319-337).  ``r2`` may be arbitrary points; pass ``grid=`` (a
    ``Grid`` whose ``coords`` are ``r2``) to use the tiled pixel-grid kernel instead."""
    import torch
    lib = L.load()
    args = (amp, phase_offset, Q1_inv, A_, B, C_, D, e, f, r1m, theta1m, k, r2)
    kind = max(A.kind_of(a) for a in args)
    dev = A.cuda_device_of(args) or torch.device("cuda", A.current_device_index())
    t = _flat_inputs(dev, amp, phase_offset, Q1_inv, A_, B, C_, D, e, f, r1m, theta1m, k)
    nb = t[0].reshape(-1).shape[0]
    poly = torch.empty((nb, 12), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.tg_beamlet_coeffs_f64(nb, *[x.data_ptr() for x in t], poly.data_ptr(),
                                          A.current_stream_ptr(dev)), "tg_beamlet_coeffs_f64")
    return _points_or_grid(poly, nb, r2, dev, grid, cull_bits, kind)


def evaluate_misaligned_input_gaussian_jax_scan(amp, phase_offset, Q1_inv, r1m, theta1m, k, r1,
                                                batch_size=128, *, grid=None, cull_bits=None):
    """Flat-array input-plane field sum (gaussian.py:410-427)."""
    import torch
    lib = L.load()
    args = (amp, phase_offset, Q1_inv, r1m, theta1m, k, r1)
    kind = max(A.kind_of(a) for a in args)
    dev = A.cuda_device_of(args) or torch.device("cuda", A.current_device_index())
    t = _flat_inputs(dev, amp, phase_offset, Q1_inv, r1m, theta1m, k)
    nb = t[0].reshape(-1).shape[0]
    poly = torch.empty((nb, 12), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.tg_input_coeffs_f64(nb, *[x.data_ptr() for x in t], poly.data_ptr(),
                                        A.current_stream_ptr(dev)), "tg_input_coeffs_f64")
    return _points_or_grid(poly, nb, r1, dev, grid, cull_bits, kind)
