"""Optical components (reference ``src/temgym_core/components.py``).

Plain frozen dataclasses with the reference's field names, order and defaults.  A component
carries parameters only; its ray arithmetic lives in the CUDA ray kernel
(``csrc/trace.cu``), selected through the opcode returned by ``_tg_spec``.  Calling a
component on a ray launches that kernel for the single step (no free-space insertion),
like the reference's ``component(ray)``.
"""
from __future__ import annotations

from dataclasses import dataclass, field, fields
from typing import Any, NamedTuple

import numpy as np

from . import Degrees
from . import _lib as L
from .aberrations import KrivanekCoeffs
from .grid import Grid
from .tree_utils import HasParamsMixin


def _f(v) -> float:
    """Component parameters are scalars (the model descriptor lives in kernel-parameter
    constant memory).  Scanner / Descanner also accept ARRAYS over the ray batch (what ``jax.vmap`` over scan
    positions gives the reference): see ``_is_array`` / ``_tg_perray``."""
    try:
        return float(v)
    except Exception as exc:  # noqa: BLE001
        raise TypeError(f"component parameter must be a scalar, got {type(v).__name__}") from exc


def _is_array(v) -> bool:
    """True for a numpy array / torch tensor / list with more than one element."""
    if isinstance(v, (list, tuple)):
        return len(v) > 1
    n = getattr(v, "numel", None)
    if callable(n):
        return n() > 1
    return getattr(v, "size", 1) > 1 and getattr(v, "ndim", 0) > 0


def _offset_spec(offsets):
    """(spec params, per-ray arrays) of a TG_OP_OFFSET component from its four offsets, scalars or arrays."""
    scal = tuple(0.0 if _is_array(o) else _f(o) for o in offsets)
    arrs = [o if _is_array(o) else None for o in offsets]
    return scal, (arrs if any(a is not None for a in arrs) else None)


class Component(HasParamsMixin):
    """Base component (components.py:12-24)."""

    def _tg_spec(self):
        """-> (opcode, z, params tuple) for the model descriptor."""
        raise NotImplementedError(
            f"{type(self).__name__} has no CUDA implementation; only the components of the "
            "accelerated path can be used in a model (no Python fallback)")

    def __call__(self, ray):
        from .run import _apply_component_only
        return _apply_component_only(ray, self)

    # parameter -> kernel tangent seeds [(slot, weight)], slot 0 = z, slot k+1 = p[k]
    # (include/temgym_b200.h, tg_trace_grad_f64).  Fields that do not influence the ray map to [].
    _TG_SLOTS = {}

    def _tg_param_seeds(self, path):
        name = path[0]
        if len(path) == 1 and name == "z":
            return [(0, 1.0)]
        if len(path) == 1 and name in self._TG_SLOTS:
            return [(self._TG_SLOTS[name], 1.0)]
        if len(path) == 1 and name in {f.name for f in fields(self)}:
            return []
        raise RuntimeError(f"Cannot find {path} in parameters of {type(self).__name__}")


class DescanError(NamedTuple):  # components.py:27-115
    pxo_pxi: float = 0.0
    pxo_pyi: float = 0.0
    pyo_pxi: float = 0.0
    pyo_pyi: float = 0.0
    sxo_pxi: float = 0.0
    sxo_pyi: float = 0.0
    syo_pxi: float = 0.0
    syo_pyi: float = 0.0
    offpxi: float = 0.0
    offpyi: float = 0.0
    offsxi: float = 0.0
    offsyi: float = 0.0

    def as_array(self) -> np.ndarray:
        return np.array(self)

    def as_matrix(self) -> np.ndarray:
        # components.py:105-115, including the reference's offsyi-twice quirk (row 3)
        return np.array([
            [self.pxo_pxi, self.pxo_pyi, 0.0, 0.0, self.offpxi],
            [self.pyo_pxi, self.pyo_pyi, 0.0, 0.0, self.offpyi],
            [self.sxo_pxi, self.sxo_pyi, 0.0, 0.0, self.offsyi],
            [self.syo_pxi, self.syo_pyi, 0.0, 0.0, self.offsyi],
            [0.0, 0.0, 0.0, 0.0, 1.0],
        ])


@dataclass(frozen=True)
class Plane(Component):  # components.py:118-134
    z: float

    def _tg_spec(self):
        return L.TG_OP_PLANE, _f(self.z), ()


@dataclass(frozen=True)
class Lens(Component):  # components.py:137-174
    z: float
    focal_length: float
    _TG_SLOTS = {"focal_length": 1}

    def _tg_spec(self):
        return L.TG_OP_LENS, _f(self.z), (_f(self.focal_length),)


@dataclass(frozen=True)
class AberratedLensKrivanek(Lens):  # components.py:177-215
    coeffs: Any = field(default_factory=KrivanekCoeffs)

    def _tg_spec(self):
        from .aberrations import krivanek_param_block
        return (L.TG_OP_KRIVANEK, _f(self.z), (_f(self.focal_length),) + krivanek_param_block(self.coeffs))

    def _tg_param_seeds(self, path):
        if path[0] == "coeffs":
            # slot 1 = focal_length, slots 2..26 = the 25 KrivanekCoeffs fields in field order
            names = [f.name for f in fields(type(self.coeffs))] if not isinstance(self.coeffs, dict) else None
            if names is None:
                from .aberrations import KrivanekCoeffs
                names = [f.name for f in fields(KrivanekCoeffs)]
            if len(path) == 2 and path[1] in names:
                return [(2 + names.index(path[1]), 1.0)]
            raise RuntimeError(f"Cannot find {path} in parameters of AberratedLensKrivanek")
        return super()._tg_param_seeds(path)


@dataclass(frozen=True)
class ScanGrid(Component, Grid):  # components.py:218-249
    z: float
    pixel_size: Any
    shape: Any
    rotation: Degrees = 0.
    centre: Any = (0., 0)
    flip_y: bool = False

    def _tg_spec(self):
        return L.TG_OP_PLANE, _f(self.z), ()


@dataclass(frozen=True)
class Scanner(Component):  # components.py:252-285
    z: float
    scan_pos_x: float
    scan_pos_y: float
    scan_tilt_x: float = 0.
    scan_tilt_y: float = 0.
    _TG_SLOTS = {"scan_pos_x": 1, "scan_pos_y": 2, "scan_tilt_x": 3, "scan_tilt_y": 4}

    def _offsets(self):
        return (self.scan_pos_x, self.scan_pos_y, self.scan_tilt_x, self.scan_tilt_y)   # components.py:279-285

    def _tg_spec(self):
        return L.TG_OP_OFFSET, _f(self.z), _offset_spec(self._offsets())[0]

    def _tg_perray(self):
        """Per-ray offset arrays (or None) when a scan position / tilt is an array over the ray batch."""
        return _offset_spec(self._offsets())[1]


@dataclass(frozen=True)
class Descanner(Component):  # components.py:288-372
    z: float
    scan_pos_x: float
    scan_pos_y: float
    scan_tilt_x: float = 0.
    scan_tilt_y: float = 0.
    descan_error: DescanError = DescanError()

    def _offsets(self):
        de = self.descan_error
        sp_x, sp_y = self.scan_pos_x, self.scan_pos_y
        st_x, st_y = self.scan_tilt_x, self.scan_tilt_y
        if not any(_is_array(v) for v in (sp_x, sp_y, st_x, st_y) + tuple(de)):
            sp_x, sp_y, st_x, st_y = _f(sp_x), _f(sp_y), _f(st_x), _f(st_y)
            de = DescanError(*(_f(v) for v in de))
        # the four 5th-column offsets, same operation order as components.py:343-372 (elementwise on arrays)
        return (sp_x * de.pxo_pxi + sp_y * de.pxo_pyi + de.offpxi - sp_x,
                sp_x * de.pyo_pxi + sp_y * de.pyo_pyi + de.offpyi - sp_y,
                sp_x * de.sxo_pxi + sp_y * de.sxo_pyi + de.offsxi - st_x,
                sp_x * de.syo_pxi + sp_y * de.syo_pyi + de.offsyi - st_y)

    def _tg_spec(self):
        return L.TG_OP_OFFSET, _f(self.z), _offset_spec(self._offsets())[0]

    def _tg_perray(self):
        return _offset_spec(self._offsets())[1]

    def _tg_param_seeds(self, path):
        # the kernel sees the four offsets o1..o4 (slots 1..4); chain rule of components.py:343-372
        de = self.descan_error
        sp_x, sp_y = _f(self.scan_pos_x), _f(self.scan_pos_y)
        if len(path) == 1:
            n = path[0]
            if n == "scan_pos_x":
                return [(1, _f(de.pxo_pxi) - 1.0), (2, _f(de.pyo_pxi)), (3, _f(de.sxo_pxi)), (4, _f(de.syo_pxi))]
            if n == "scan_pos_y":
                return [(1, _f(de.pxo_pyi)), (2, _f(de.pyo_pyi) - 1.0), (3, _f(de.sxo_pyi)), (4, _f(de.syo_pyi))]
            if n == "scan_tilt_x":
                return [(3, -1.0)]
            if n == "scan_tilt_y":
                return [(4, -1.0)]
            return super()._tg_param_seeds(path)
        if len(path) == 2 and path[0] == "descan_error":
            n = path[1]
            if isinstance(n, int):
                n = DescanError._fields[n]
            table = {"pxo_pxi": (1, sp_x), "pxo_pyi": (1, sp_y), "pyo_pxi": (2, sp_x), "pyo_pyi": (2, sp_y),
                     "sxo_pxi": (3, sp_x), "sxo_pyi": (3, sp_y), "syo_pxi": (4, sp_x), "syo_pyi": (4, sp_y),
                     "offpxi": (1, 1.0), "offpyi": (2, 1.0), "offsxi": (3, 1.0), "offsyi": (4, 1.0)}
            if n in table:
                return [table[n]]
        raise RuntimeError(f"Cannot find {path} in parameters of Descanner")


@dataclass(frozen=True)
class Detector(Component, Grid):  # components.py:375-406
    z: float
    pixel_size: Any
    shape: Any
    rotation: Degrees = 0.
    centre: Any = (0., 0)
    flip_y: bool = False

    def _tg_spec(self):
        return L.TG_OP_PLANE, _f(self.z), ()


@dataclass(frozen=True)
class ThickLens(Component):  # components.py:409-452
    z_po: float
    z_pi: float
    focal_length: float

    @property
    def z(self):
        return self.z_po

    def _tg_spec(self):
        return L.TG_OP_THICKLENS, _f(self.z_po), (_f(self.focal_length), _f(self.z_po) - _f(self.z_pi))

    def _tg_param_seeds(self, path):
        n = path[0]
        if n == "focal_length":
            return [(1, 1.0)]
        if n == "z_po":      # plane position and the z jump p[1] = z_po - z_pi
            return [(0, 1.0), (2, 1.0)]
        if n == "z_pi":
            return [(2, -1.0)]
        raise RuntimeError(f"Cannot find {path} in parameters of ThickLens")


@dataclass(frozen=True)
class Deflector(Component):  # components.py:455-482
    z: float
    def_x: float
    def_y: float
    _TG_SLOTS = {"def_x": 1, "def_y": 2}

    def _tg_spec(self):
        return L.TG_OP_DEFLECTOR, _f(self.z), (_f(self.def_x), _f(self.def_y))


@dataclass(frozen=True)
class Rotator(Component):  # components.py:485-523
    z: float
    angle: Degrees

    def _tg_spec(self):
        a = np.deg2rad(_f(self.angle))
        return L.TG_OP_ROTATOR, _f(self.z), (float(np.cos(a)), float(np.sin(a)))

    def _tg_param_seeds(self, path):
        if path[0] == "angle":  # p = (cos a, sin a), a = angle * pi / 180
            a = np.deg2rad(_f(self.angle))
            k = np.pi / 180.0
            return [(1, float(-np.sin(a) * k)), (2, float(np.cos(a) * k))]
        return super()._tg_param_seeds(path)


@dataclass(frozen=True)
class Biprism(Component):  # components.py:526-559 (offset, rotation, side are unused there too)
    z: float
    offset: float = 0.0
    rotation: Degrees = 0.0
    def_x: float = 0.0
    side: int = 1
    _TG_SLOTS = {"def_x": 1}

    def _tg_spec(self):
        return L.TG_OP_BIPRISM, _f(self.z), (_f(self.def_x),)
