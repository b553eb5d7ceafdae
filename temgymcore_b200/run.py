"""Runners (reference ``src/temgym_core/run.py``): ``run_iter``, ``run_to_end``,
``solve_model`` -- and the fused ray + ABCD call that replaces
``jax.vmap(jax.jacobian(run_to_end), in_axes=(0, None))`` + ``custom_jacobian_matrix``
(reference gaussian.py:234-239, utils.py:7-43).

A model (any sequence of components / sources) is compiled to a flat descriptor
(opcode, z, parameters per element) that travels to the CUDA ray kernel as a kernel
parameter; one launch propagates every ray through the whole model, carrying the
Jacobian with forward-mode duals in registers.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Generator, Sequence

import numpy as np

from . import _arrays as A
from . import _lib as L
from .components import Component
from .propagator import BasePropagator, FreeSpaceParaxial, Propagator
from .ray import RAY_FIELDS, Ray
from .source import Source


# ----------------------------------------------------------------------------- model compile
def compile_model(components: Sequence[Any], noprop: bool = False) -> L.tg_model:
    """Flatten a model into the ``tg_model`` descriptor of ``include/temgym_b200.h``."""
    comps = list(components)
    if len(comps) > L.TG_MAX_COMPS:
        raise ValueError(f"model has {len(comps)} elements; the kernel descriptor holds "
                         f"{L.TG_MAX_COMPS}")
    m = L.tg_model()
    m.n_comp = len(comps)
    perray = []
    for i, c in enumerate(comps):
        if isinstance(c, Source):
            op, z, params = L.TG_OP_PLANE, float(c.z), ()
        elif isinstance(c, Component):
            op, z, params = c._tg_spec()
        else:
            raise TypeError(f"model element {i} ({type(c).__name__}) is neither a Component nor a "
                            "Source")
        m.comp[i].op = op
        m.comp[i].flags = L.TG_F_NOPROP if noprop else 0
        m.comp[i].z = z
        for j, v in enumerate(params):
            m.comp[i].p[j] = v
        arrs = c._tg_perray() if hasattr(c, "_tg_perray") else None
        if arrs is not None:               # Scanner / Descanner with array-valued parameters (per-ray offsets)
            perray.append((i, arrs))
    if len(perray) > L.TG_MAX_PERRAY:
        raise ValueError(f"at most {L.TG_MAX_PERRAY} components of a model may carry per-ray parameters")
    m._perray = perray
    return m


def require_scalar_params(model: L.tg_model, what: str) -> L.tg_model:
    """Per-ray component parameters are served by the ray kernel (run_to_end / run_iter / the ABCD calls);
    everything else takes scalar-parameter models."""
    if getattr(model, "_perray", None):
        raise NotImplementedError(f"{what}: array-valued Scanner / Descanner parameters are supported by run_to_end, "
                                  "run_iter, run_to_end_abcd and ray_jacobian only")
    return model


def _host_z(z, model: L.tg_model) -> float:
    """z after the model for a ray-independent (scalar) input z: the same fp64 arithmetic
    the kernel performs (run.py:77 + propagator.py:70; ThickLens components.py:442)."""
    for i in range(model.n_comp):
        c = model.comp[i]
        if not (c.flags & L.TG_F_NOPROP):
            z = z + (c.z if (c.flags & L.TG_F_DIST) else (c.z - z))
        if c.op == L.TG_OP_THICKLENS:
            z = z - c.p[1]
    return z


def _host_empty(shape) -> np.ndarray:
    """fp64 host result buffer; page-locked (torch's caching host allocator) for large
    results so the D2H copy inside tg_*_host runs at PCIe rate."""
    n = int(np.prod(shape))
    if n >= (1 << 15):
        try:
            import torch
            return torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()
        except Exception:  # noqa: BLE001 - no CUDA runtime: the call below fails loudly anyway
            pass
    return np.empty(shape, dtype=np.float64)


# ----------------------------------------------------------------------------- kernel call
def _trace(ray, model: L.tg_model, jac_layout: int = L.TG_JAC_NONE, want_rays: bool = True):
    """Run the ray kernel.  Returns (Ray | None, jac | None)."""
    lib = L.load()
    vals = [getattr(ray, f) for f in RAY_FIELDS]
    perray = getattr(model, "_perray", None)
    if perray:
        return _trace_perray(ray, model, perray, jac_layout, want_rays)
    kinds = [A.kind_of(v) for v in vals]
    kind = max(kinds)
    sizes = [A.numel(v) for v in vals]
    n = max(sizes)
    for f, k, s in zip(RAY_FIELDS, kinds, sizes):
        if k != A.KIND_SCALAR and s != n:
            raise ValueError(f"Ray field {f!r} has {s} elements, expected {n}")
    shape = next((A.shape_of(v) for v, k in zip(vals, kinds) if k != A.KIND_SCALAR), ())
    jw = {L.TG_JAC_NONE: 0, L.TG_JAC_ABCD5: 25, L.TG_JAC_FULL7: 49}[jac_layout]
    jdim = 5 if jw == 25 else 7
    # z and _one stay ray-independent scalars when they came in as scalars
    scalar_out = {4: kinds[4] == A.KIND_SCALAR and kind != A.KIND_SCALAR,
                  6: kinds[6] == A.KIND_SCALAR and kind != A.KIND_SCALAR}
    rin = L.tg_ray_in()
    out_ptrs = [None] * 7
    outs = [None] * 7
    keep = []

    if kind == A.KIND_CUDA:
        import torch
        dev = A.cuda_device_of(vals)
        for i, (v, k) in enumerate(zip(vals, kinds)):
            if k == A.KIND_SCALAR:
                rin.ptr[i] = None
                rin.value[i] = A.to_float(v)
            else:
                t = A.to_device_f64(v, dev)
                keep.append(t)
                rin.ptr[i] = t.data_ptr()
        if want_rays:
            for i in range(7):
                if scalar_out.get(i, False):
                    continue
                outs[i] = torch.empty(n, dtype=torch.float64, device=dev)
                out_ptrs[i] = outs[i].data_ptr()
        jac = torch.empty((n, jdim, jdim), dtype=torch.float64, device=dev) if jw else None
        with torch.cuda.device(dev):
            L.check(lib.tg_trace_f64(C.byref(model), n, C.byref(rin), L.ptr_array(out_ptrs),
                                     jac.data_ptr() if jw else None, jac_layout,
                                     A.current_stream_ptr(dev)), "tg_trace_f64")
        conv = lambda t: t.reshape(shape)  # noqa: E731
        if jw:
            jac = jac.reshape(shape + (jdim, jdim))
    else:
        for i, (v, k) in enumerate(zip(vals, kinds)):
            if k == A.KIND_SCALAR:
                rin.ptr[i] = None
                rin.value[i] = A.to_float(v)
            else:
                h = A.to_host_f64(v)
                keep.append(h)
                rin.ptr[i] = h.ctypes.data
        if want_rays:
            for i in range(7):
                if scalar_out.get(i, False):
                    continue
                outs[i] = _host_empty((n,))
                out_ptrs[i] = outs[i].ctypes.data
        jac = _host_empty((n, jdim, jdim)) if jw else None
        L.check(lib.tg_trace_f64_host(C.byref(model), n, C.byref(rin), L.ptr_array(out_ptrs),
                                      jac.ctypes.data if jw else None, jac_layout,
                                      A.current_device_index()), "tg_trace_f64_host")
        if kind == A.KIND_SCALAR:
            conv = lambda a: float(a[0])  # noqa: E731
            if jw:
                jac = jac[0]
        else:
            conv = lambda a: A.from_host(a, kind, shape)  # noqa: E731
            if jw:
                jac = A.from_host(jac, kind, shape + (jdim, jdim))

    out_ray = None
    if want_rays:
        res = []
        for i in range(7):
            if outs[i] is not None:
                res.append(conv(outs[i]))
            elif i == 4:
                res.append(_host_z(A.to_float(vals[4]), model))
            else:  # _one: `one * 1.0`
                res.append(A.to_float(vals[6]) * 1.0)
        out_ray = Ray(*res)
    return out_ray, jac


def _trace_perray(ray, model, perray, jac_layout, want_rays):
    """Ray kernel with per-ray Scanner / Descanner offsets (``tg_trace_perray_f64``): the batch is the common
    length of the ray arrays and the parameter arrays (scalars broadcast), everything runs on the device and
    the results come back in the kind of the inputs (CUDA tensors stay, numpy / CPU tensors are copied back)."""
    import torch
    lib = L.load()
    vals = [getattr(ray, f) for f in RAY_FIELDS]
    arrays = [a for _, arrs in perray for a in arrs if a is not None]
    kind = max([A.kind_of(v) for v in vals] + [A.kind_of(a) for a in arrays])
    n = max([A.numel(v) for v in vals] + [A.numel(a) for a in arrays])
    shape = next((A.shape_of(v) for v in vals + arrays if A.numel(v) == n), (n,))
    for v in vals + arrays:
        if A.numel(v) not in (1, n):
            raise ValueError(f"ray fields and per-ray component parameters must share one batch size ({n})")
    dev = A.cuda_device_of(vals + arrays) or torch.device("cuda", A.current_device_index())
    rin = L.tg_ray_in()
    keep = []
    for i, v in enumerate(vals):
        if A.numel(v) == 1:
            rin.ptr[i] = None
            rin.value[i] = A.to_float(v)
        else:
            t = A.to_device_f64(v, dev)
            keep.append(t)
            rin.ptr[i] = t.data_ptr()
    pr = L.tg_perray()
    pr.n = len(perray)
    for k, (ci, arrs) in enumerate(perray):
        pr.comp[k] = ci
        for j, a in enumerate(arrs):
            if a is None:
                pr.ptr[k][j] = None
            else:
                t = A.to_device_f64(a, dev)
                t = t.expand(n).contiguous() if t.numel() == 1 else t
                keep.append(t)
                pr.ptr[k][j] = t.data_ptr()
    jw = {L.TG_JAC_NONE: 0, L.TG_JAC_ABCD5: 25, L.TG_JAC_FULL7: 49}[jac_layout]
    jdim = 5 if jw == 25 else 7
    outs = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(7)] if want_rays else None
    jac = torch.empty((n, jdim, jdim), dtype=torch.float64, device=dev) if jw else None
    with torch.cuda.device(dev):
        L.check(lib.tg_trace_perray_f64(C.byref(model), n, C.byref(rin), C.byref(pr),
                                        L.ptr_array([o.data_ptr() for o in outs] if outs else [None] * 7),
                                        jac.data_ptr() if jw else None, jac_layout, A.current_stream_ptr(dev)),
                "tg_trace_perray_f64")

    def back(t, shp):
        t = t.reshape(shp)
        if kind == A.KIND_CUDA:
            return t
        return t.cpu() if kind == A.KIND_TORCH_CPU else t.cpu().numpy()
    out_ray = Ray(*(back(o, shape) for o in outs)) if want_rays else None
    return out_ray, (back(jac, tuple(shape) + (jdim, jdim)) if jw else None)


def _check_propagator(propagator):
    if not isinstance(propagator, FreeSpaceParaxial):
        raise NotImplementedError(
            "only FreeSpaceParaxial is implemented by the CUDA ray kernel "
            f"(got {type(propagator).__name__}); there is no Python fallback")


def _distance_model(distance) -> L.tg_model:
    if A.numel(distance) != 1:
        raise NotImplementedError("per-ray propagation distances: use run_to_end with a model")
    m = L.tg_model()
    m.n_comp = 1
    m.comp[0].op = L.TG_OP_PLANE
    m.comp[0].flags = L.TG_F_DIST
    m.comp[0].z = A.to_float(distance)
    return m


def _propagate_only(ray, distance):
    """FreeSpaceParaxial.propagate(ray, distance) (propagator.py:52-72) on the GPU."""
    out, _ = _trace(ray, _distance_model(distance))
    return out


def _apply_component_only(ray, component):
    """``component(ray)`` with no free-space step (components.py ``__call__``)."""
    out, _ = _trace(ray, compile_model([component], noprop=True))
    return out


# ----------------------------------------------------------------------------- public API
def passthrough_transform(component):  # run.py:26-32
    def inner(ray):
        out = component(ray)
        return out, out
    return inner


def jacobian_transform(component):  # run.py:35-41
    def inner(ray):
        from .utils import RayJacobian
        if isinstance(component, Propagator):
            m = _distance_model(component.distance)
        else:
            m = compile_model([component], noprop=True)
        out, jac = _trace(ray, m, L.TG_JAC_FULL7)
        return out, RayJacobian(jac)
    return inner


def run_iter(ray, components: Sequence[Any], transform=passthrough_transform,
             propagator: BasePropagator = FreeSpaceParaxial()) -> Generator:
    """Step a ray through the model, yielding ``(Propagator, ray)`` then
    ``(component, ray)`` for every element; free space over ``component.z - ray.z`` is
    inserted before every element, also when that distance is 0 (run.py:47-82)."""
    _check_propagator(propagator)
    for component in components:
        if isinstance(component, (Source, Component)):
            distance = component.z - ray.z
            propagator_d = propagator.with_distance(distance)
            if transform is passthrough_transform:
                m = L.tg_model()
                m.n_comp = 1
                m.comp[0].op = L.TG_OP_PLANE
                m.comp[0].z = float(component.z)
                ray, _ = _trace(ray, m)
                out = ray
            else:
                ray, out = transform(propagator_d)(ray)
            yield propagator_d, out
        ray, out = transform(component)(ray)
        yield component, out


def run_to_end(ray, components: Sequence[Any],
               propagator: BasePropagator = FreeSpaceParaxial()) -> Ray:
    """Propagate ray(s) through all components in ONE kernel launch (run.py:85-116)."""
    _check_propagator(propagator)
    out, _ = _trace(ray, compile_model(components))
    return out


def run_to_end_abcd(ray, components: Sequence[Any],
                    propagator: BasePropagator = FreeSpaceParaxial(), want_rays: bool = True):
    """Fused ``run_to_end`` + per-ray 5x5 ABCD matrix, one launch.

    Replaces ``custom_jacobian_matrix(jax.vmap(jax.jacobian(run_to_end), in_axes=(0, None))
    (rays, model))`` (gaussian.py:234-239).  Returns ``(Ray, abcd)`` with ``abcd`` of shape
    ``(N, 5, 5)`` (``(5, 5)`` for a scalar ray)."""
    _check_propagator(propagator)
    return _trace(ray, compile_model(components), L.TG_JAC_ABCD5, want_rays=want_rays)


class RayTracePlan:
    """``run_to_end`` (+ ABCD) for a fixed model and ray count, captured once in a CUDA graph.

    A C1-sized launch (1e6 rays) is a ~55 us kernel; building the descriptor, allocating the
    results and crossing ctypes costs the host more than that per call.  The plan compiles the
    model once, owns static input / output buffers and replays one graph node per ``run()``:
    ``update(ray)`` copies new rays (same count) in, ``run()`` returns ``(Ray, abcd)`` views of
    the static outputs (``abcd`` is None when ``jacobian`` is False).  The model is baked in.
    """

    def __init__(self, ray, components: Sequence[Any], *, jacobian: bool = True, device=None):
        import torch
        vals = [getattr(ray, f) for f in RAY_FIELDS]
        dev = A.cuda_device_of(vals) or torch.device("cuda", A.current_device_index()
                                                     if device is None else device)
        self.device = dev
        self._layout = L.TG_JAC_ABCD5 if jacobian else L.TG_JAC_NONE
        self._model = require_scalar_params(compile_model(components), "RayTracePlan")
        self._static = Ray(*(v if A.kind_of(v) == A.KIND_SCALAR else A.to_device_f64(v, dev).clone()
                             for v in vals))
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                      # warm-up outside the capture
            _trace(self._static, self._model, self._layout)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._out = _trace(self._static, self._model, self._layout)

    def update(self, ray):
        for f in RAY_FIELDS:
            dst, src = getattr(self._static, f), getattr(ray, f)
            if A.kind_of(dst) == A.KIND_SCALAR:
                if A.to_float(src) != dst:
                    raise ValueError(f"scalar ray field {f!r} is baked into the plan")
                continue
            dst.copy_(A.to_device_f64(src, self.device), non_blocking=True)
        return self

    def run(self):
        self._graph.replay()
        return self._out

    __call__ = run


def ray_jacobian(ray, components: Sequence[Any], propagator: BasePropagator = FreeSpaceParaxial()):
    """``jax.jacobian(run_to_end)(ray, model)`` (README.md:227-234): the full Ray-of-Ray
    Jacobian as a :class:`~temgymcore_b200.utils.RayJacobian` (attribute access
    ``jac.dy.x`` == d out.dy / d in.x)."""
    from .utils import RayJacobian
    _check_propagator(propagator)
    _, jac = _trace(ray, compile_model(components), L.TG_JAC_FULL7, want_rays=False)
    return RayJacobian(jac)


def calculate_derivatives(ray, model: Sequence[Any], order: int):
    """Successive forward-mode derivatives of ``run_to_end`` w.r.t. the ray (run.py:119-147).

    Returns a list of ``order`` :class:`~temgymcore_b200.utils.RayDerivative` objects; entry ``k-1``
    holds the dense tensor ``d^k out_f / d in_a1 ... d in_ak`` (shape ``batch + (7,) * (k + 1)``) that
    ``k`` nested ``jax.jacfwd(run_to_end, argnums=0)`` calls produce, with the reference's nested
    attribute access (``derivs[1].x.dx.dy``).  One launch of the hyper-dual CUDA kernel
    (``tg_trace_jets_f64``) serves all orders up to 3; higher orders are not implemented."""
    from .utils import RayDerivative
    import torch
    order = int(order)
    if order < 1:
        return []
    if order > 3:
        raise NotImplementedError("calculate_derivatives: the CUDA jet kernel implements orders 1..3")
    lib = L.load()
    cm = require_scalar_params(compile_model(model), "calculate_derivatives")
    vals = [getattr(ray, f) for f in RAY_FIELDS]
    kinds = [A.kind_of(v) for v in vals]
    kind = max(kinds)
    n = max(A.numel(v) for v in vals)
    shape = next((A.shape_of(v) for v, k in zip(vals, kinds) if k != A.KIND_SCALAR), ())
    dev = A.cuda_device_of(vals) or torch.device("cuda", A.current_device_index())
    rin = L.tg_ray_in()
    keep = []
    for i, (v, k) in enumerate(zip(vals, kinds)):
        if k == A.KIND_SCALAR:
            rin.ptr[i] = None
            rin.value[i] = A.to_float(v)
        else:
            t = A.to_device_f64(v, dev)
            keep.append(t)
            rin.ptr[i] = t.data_ptr()
    tensors = [torch.empty((n,) + (7,) * (k + 1), dtype=torch.float64, device=dev) for k in range(1, order + 1)]
    ptrs = [t.data_ptr() for t in tensors] + [None] * (3 - order)
    with torch.cuda.device(dev):
        L.check(lib.tg_trace_jets_f64(C.byref(cm), n, C.byref(rin), order, L.ptr_array([None] * 7),
                                      ptrs[0], ptrs[1], ptrs[2], A.current_stream_ptr(dev)),
                "tg_trace_jets_f64")

    def finish(t, k):
        t = t.reshape(shape + (7,) * (k + 1))
        if kind == A.KIND_CUDA:
            return t
        return t.cpu() if kind == A.KIND_TORCH_CPU else t.cpu().numpy()

    return [RayDerivative(finish(t, k), k) for k, t in enumerate(tensors, start=1)]


def solve_model(ray, model: Sequence[Any], propagator: BasePropagator = FreeSpaceParaxial()):
    """Per-step 5x5 ABCD matrices, shape ``(2 * n_components, 5, 5)`` (run.py:150-179)."""
    from .utils import custom_jacobian_matrix
    mats = []
    for _, jac in run_iter(ray, model, transform=jacobian_transform, propagator=propagator):
        mats.append(custom_jacobian_matrix(jac))
    if A.kind_of(mats[0]) == A.KIND_CUDA:
        import torch
        return torch.stack(mats)
    return np.array(mats)


def _expand_param_leaves(root, path):
    """``[(key_suffix, leaf_path)]`` for the parameter of ``root`` at ``path``: a scalar leaf is itself
    (empty suffix); a NamedTuple / dataclass node expands into its fields in field order, keyed by index."""
    import dataclasses
    try:
        v = root
        for k in path:
            v = getattr(v, k) if isinstance(k, str) else v[k]
    except (AttributeError, IndexError, KeyError, TypeError):
        return [((), path)]                 # let _tg_param_seeds raise the reference's RuntimeError
    if isinstance(v, tuple) and hasattr(v, "_fields"):
        return [((i,), path + (name,)) for i, name in enumerate(v._fields)]
    if dataclasses.is_dataclass(v) and not isinstance(v, type):
        return [((i,), path + (f.name,)) for i, f in enumerate(dataclasses.fields(v))]
    return [((), path)]


def run_with_grads(input_ray, model: Sequence[Any], grad_vars: Sequence[Any]):
    """Run the model and compute Jacobians w.r.t. selected variables (run.py:182-267).

    ``grad_vars`` holds ``input_ray`` itself (all seven fields), ``input_ray.params.<field>`` or
    ``component.params.<name>`` references (components of ``model`` are matched by identity).
    Returns ``(value, grads)``: the output ``Ray`` and a dict mapping each variable's path
    ``(root, key, ...)`` to a ``Ray``-shaped Jacobian (d out_field / d variable per ray).
    One launch of the CUDA gradient kernel serves up to 8 variables.
    """
    from .tree_utils import ParamRef
    lib = L.load()
    comps = list(model)
    directions = []          # (key, ray_field_index | None, [(comp, slot, weight)])
    for var in grad_vars:
        if var is input_ray:
            for i, f in enumerate(RAY_FIELDS):
                directions.append(((input_ray, f), i, []))
            continue
        if not isinstance(var, ParamRef):
            raise RuntimeError(f"Cannot find {var!r} in parameters")
        root, path = var._resolve_root(), var._build()[1:]
        if root is input_ray:
            if len(path) != 1 or path[0] not in RAY_FIELDS:
                raise RuntimeError(f"Cannot find {var._build()} in parameters")
            directions.append((var._build(), RAY_FIELDS.index(path[0]), []))
            continue
        idx = next((i for i, c in enumerate(comps) if c is root), None)
        if idx is None or not path:
            raise RuntimeError(f"Cannot find {var._build()} in parameters")
        # a reference to a container-valued parameter (DescanError, KrivanekCoeffs) stands for all of its
        # leaves: the reference expands the node into one variable per leaf, keyed path + (leaf index,)
        # (PathBuilder._find_in, tree_utils.py:100-122)
        for suffix, leaf in _expand_param_leaves(root, tuple(path)):
            seeds = [(idx, slot, w) for slot, w in root._tg_param_seeds(leaf)]
            directions.append((var._build() + suffix, None, seeds))
    if not directions:
        raise RuntimeError("Cannot find any variable in parameters")

    import torch
    cm = require_scalar_params(compile_model(comps), "run_with_grads")
    vals = [getattr(input_ray, f) for f in RAY_FIELDS]
    kinds = [A.kind_of(v) for v in vals]
    kind = max(kinds)
    n = max(A.numel(v) for v in vals)
    shape = next((A.shape_of(v) for v, k in zip(vals, kinds) if k != A.KIND_SCALAR), ())
    dev = A.cuda_device_of(vals) or torch.device("cuda", A.current_device_index())
    rin = L.tg_ray_in()
    keep = []
    for i, (v, k) in enumerate(zip(vals, kinds)):
        if k == A.KIND_SCALAR:
            rin.ptr[i] = None
            rin.value[i] = A.to_float(v)
        else:
            t = A.to_device_f64(v, dev)
            keep.append(t)
            rin.ptr[i] = t.data_ptr()
    outs = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(7)]
    lanes = L.TG_GRAD_LANES
    grads = {}

    def finish(t):
        t = t.reshape(shape)
        if kind == A.KIND_CUDA:
            return t
        if kind == A.KIND_SCALAR:
            return float(t.reshape(-1)[0].item())
        return t.cpu() if kind == A.KIND_TORCH_CPU else t.cpu().numpy()

    for g0 in range(0, len(directions), lanes):
        group = directions[g0:g0 + lanes]
        ray_lane = (C.c_int32 * 7)(*([-1] * 7))
        seeds = []
        for lane, (_, rf, sd) in enumerate(group):
            if rf is not None:
                if ray_lane[rf] != -1:
                    raise RuntimeError("a ray field was requested twice")
                ray_lane[rf] = lane
            seeds += [(c, slot, lane, w) for c, slot, w in sd]
        if len(seeds) > L.TG_MAX_SEEDS:
            raise RuntimeError("too many parameter seeds in one group")
        arr = (L.tg_seed * max(1, len(seeds)))()
        for i, (c, slot, lane, w) in enumerate(seeds):
            arr[i].comp, arr[i].slot, arr[i].lane, arr[i].weight = c, slot, lane, w
        jac = torch.empty((n, 7, lanes), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            L.check(lib.tg_trace_grad_f64(C.byref(cm), n, C.byref(rin), ray_lane, arr, len(seeds),
                                          L.ptr_array([o.data_ptr() for o in outs]), jac.data_ptr(),
                                          A.current_stream_ptr(dev)), "tg_trace_grad_f64")
        for lane, (key, _, _) in enumerate(group):
            grads[key] = Ray(*(finish(jac[:, r, lane]) for r in range(7)))
    value = Ray(*(finish(o) for o in outs))
    return value, grads
