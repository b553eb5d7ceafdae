"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the TemGymCore hot path.

A numpy fp64 / complex128 restatement of the reference algorithms.  The
reference (``/root/reference``, pure Python on JAX) cannot be imported in this
environment because ``jax`` / ``jax_dataclasses`` are not installed
(``src/temgym_core/__init__.py:4``), so this file restates the arithmetic of
the reference, function by function, citing the reference ``file:line`` each
function follows.  It is pinned by the reference's own golden vectors (see
``tests/golden/reference_goldens.json`` and ``tests/test_oracle_goldens.py``).

Pinned (against reference goldens): ray propagation, 5x5 ABCD Jacobians,
``Qinv_ABCD``, the free-space field-sum KAT (rtol 1e-9), grid pixel<->metre
tables.  **Parity unpinned** (no tight reference golden exists, SURVEY.md
section 8c): misaligned/astigmatic beamlets through non-trivial ABCDs,
``AberratedLensKrivanek`` Jacobians.  For those the oracle is a restatement only.

Nothing in the product package (``temgymcore_b200``) imports this module.

Components are duck-typed by class name (``type(c).__name__``) and attribute
names identical to the reference dataclasses, so tests can hand the *same*
model objects to the product and to the oracle without the oracle importing
the product.

State vector order everywhere: ``[x, y, dx, dy, z, pathlength, _one]``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, fields, replace
from typing import Any, Iterable, Sequence

import numpy as np

# --------------------------------------------------------------------------
# forward-mode dual numbers (restates jax.jacobian for these smooth ops; the
# reference uses reverse mode, run.py:17-41 / gaussian.py:234 -- identical
# values for the elementwise ops on this path; sign() has zero derivative).
# --------------------------------------------------------------------------


class Dual:
    """value ``v`` of shape (N,) with tangents ``t`` of shape (N, K)."""

    __slots__ = ("v", "t")
    __array_priority__ = 1000

    def __init__(self, v, t):
        self.v = np.asarray(v, dtype=np.float64)
        self.t = np.asarray(t, dtype=np.float64)

    @staticmethod
    def lift(x, like: "Dual") -> "Dual":
        if isinstance(x, Dual):
            return x
        v = np.broadcast_to(np.asarray(x, dtype=np.float64), like.v.shape)
        return Dual(v, np.zeros(like.t.shape))

    def __add__(self, o):
        if isinstance(o, Dual):
            return Dual(self.v + o.v, self.t + o.t)
        return Dual(self.v + o, self.t)

    __radd__ = __add__

    def __sub__(self, o):
        if isinstance(o, Dual):
            return Dual(self.v - o.v, self.t - o.t)
        return Dual(self.v - o, self.t)

    def __rsub__(self, o):
        return Dual(o - self.v, -self.t)

    def __neg__(self):
        return Dual(-self.v, -self.t)

    def __mul__(self, o):
        if isinstance(o, Dual):
            return Dual(self.v * o.v, self.t * o.v[..., None] + o.t * self.v[..., None])
        o = np.asarray(o, dtype=np.float64)
        return Dual(self.v * o, self.t * (o[..., None] if o.ndim else o))

    __rmul__ = __mul__

    def __truediv__(self, o):
        if isinstance(o, Dual):
            q = self.v / o.v
            return Dual(q, (self.t - o.t * q[..., None]) / o.v[..., None])
        o = np.asarray(o, dtype=np.float64)
        return Dual(self.v / o, self.t / (o[..., None] if o.ndim else o))

    def __rtruediv__(self, o):
        q = o / self.v
        return Dual(q, -self.t * (q / self.v)[..., None])

    def __pow__(self, p):
        if p == 2:
            return self * self
        return Dual(self.v ** p, self.t * (p * self.v ** (p - 1))[..., None])


# --------------------------------------------------------------------------
# truncated multivariate Taylor polynomials ("jets") -- restates nested jax.jacfwd
# (run.py:119-147, calculate_derivatives) for derivative orders <= 3.  A Jet holds the Taylor
# coefficients c_i (f(v0 + h) = sum_i c_i h^i, |i| <= order) of a quantity in the 7 input-ray
# variables; d^i f = i! c_i.  Deliberately a different algorithm from the CUDA kernel's hyper-dual
# numbers (and it follows the reference's hypot / arctan2 / cos / sin formulas literally).
# --------------------------------------------------------------------------


class JetSpace:
    """Monomial bookkeeping for ``nvar`` variables up to total degree ``order``."""

    def __init__(self, nvar: int, order: int):
        self.nvar, self.order = nvar, order
        monos = [()]
        for deg in range(1, order + 1):
            monos += [m for m in _multisets(nvar, deg)]
        # a monomial is the sorted tuple of its variable indices (with repetition)
        self.monos = monos
        self.index = {m: k for k, m in enumerate(monos)}
        pairs = []
        for i, a in enumerate(monos):
            for j, b in enumerate(monos):
                if len(a) + len(b) <= order:
                    pairs.append((i, j, self.index[tuple(sorted(a + b))]))
        self.pairs = pairs


def _multisets(nvar, deg, start=0):
    if deg == 0:
        yield ()
        return
    for v in range(start, nvar):
        for rest in _multisets(nvar, deg - 1, v):
            yield (v,) + rest


class Jet:
    """Taylor coefficients ``c`` of shape (N, n_monomials) in ``space``."""

    __slots__ = ("c", "space")
    __array_priority__ = 1000

    def __init__(self, c, space):
        self.c = np.asarray(c, dtype=np.float64)
        self.space = space

    @staticmethod
    def const(v, like: "Jet") -> "Jet":
        c = np.zeros(like.c.shape)
        c[:, 0] = v
        return Jet(c, like.space)

    @staticmethod
    def variable(v, var: int, space: JetSpace) -> "Jet":
        v = np.asarray(v, dtype=np.float64)
        c = np.zeros((v.shape[0], len(space.monos)))
        c[:, 0] = v
        if space.order >= 1:
            c[:, space.index[(var,)]] = 1.0
        return Jet(c, space)

    def _lift(self, o):
        return o if isinstance(o, Jet) else Jet.const(np.asarray(o, dtype=np.float64), self)

    def __add__(self, o):
        return Jet(self.c + self._lift(o).c, self.space)

    __radd__ = __add__

    def __sub__(self, o):
        return Jet(self.c - self._lift(o).c, self.space)

    def __rsub__(self, o):
        return Jet(self._lift(o).c - self.c, self.space)

    def __neg__(self):
        return Jet(-self.c, self.space)

    def __mul__(self, o):
        if not isinstance(o, Jet):
            o = np.asarray(o, dtype=np.float64)
            return Jet(self.c * (o[..., None] if o.ndim else o), self.space)
        out = np.zeros(self.c.shape)
        for i, j, k in self.space.pairs:
            out[:, k] += self.c[:, i] * o.c[:, j]
        return Jet(out, self.space)

    __rmul__ = __mul__

    def compose(self, taylor):
        """g(self) from the Taylor coefficients ``taylor[k] = g^(k)(a0) / k!`` at a0 = self.c[:, 0]."""
        n = Jet(self.c.copy(), self.space)
        n.c[:, 0] = 0.0
        out = Jet.const(taylor[0], self)
        pw = None
        for k in range(1, self.space.order + 1):
            pw = n if pw is None else pw * n
            out = out + pw * np.asarray(taylor[k])
        return out

    def recip(self):
        with np.errstate(divide="ignore", invalid="ignore"):
            i = 1.0 / self.c[:, 0]
            return self.compose([i, -i ** 2, i ** 3, -i ** 4])

    def sqrt(self):
        with np.errstate(divide="ignore", invalid="ignore"):
            a0 = self.c[:, 0]
            s = np.sqrt(a0)
            return self.compose([s, 0.5 / s, -0.125 / (s * a0), 0.0625 / (s * a0 * a0)])

    def __truediv__(self, o):
        if isinstance(o, Jet):
            return self * o.recip()
        o = np.asarray(o, dtype=np.float64)
        return Jet(self.c / (o[..., None] if o.ndim else o), self.space)

    def __rtruediv__(self, o):
        return self.recip() * o

    def __pow__(self, p):
        if p == 2:
            return self * self
        raise NotImplementedError

    def derivative_tensor(self, k: int) -> np.ndarray:
        """Dense symmetric (N, nvar, ..., nvar) tensor of the k-th derivatives (i! c_i)."""
        nv = self.space.nvar
        out = np.zeros((self.c.shape[0],) + (nv,) * k)
        import itertools
        import math
        for m in _multisets(nv, k):
            fact = 1
            for v in set(m):
                fact *= math.factorial(m.count(v))
            val = self.c[:, self.space.index[m]] * fact
            for perm in set(itertools.permutations(m)):
                out[(slice(None),) + perm] = val
        return out


def _sign(x):
    # jnp.sign: derivative is zero; sign(0) == 0  (components.py:556)
    if isinstance(x, Jet):
        return Jet.const(np.sign(x.c[:, 0]), x)
    if isinstance(x, Dual):
        return Dual(np.sign(x.v), np.zeros_like(x.t))
    return np.sign(x)


def _cos(x):
    if isinstance(x, Jet):
        c, sn = np.cos(x.c[:, 0]), np.sin(x.c[:, 0])
        return x.compose([c, -sn, -c / 2.0, sn / 6.0])
    if isinstance(x, Dual):
        return Dual(np.cos(x.v), -np.sin(x.v)[..., None] * x.t)
    return np.cos(x)


def _sin(x):
    if isinstance(x, Jet):
        c, sn = np.cos(x.c[:, 0]), np.sin(x.c[:, 0])
        return x.compose([sn, c, -sn / 2.0, -c / 6.0])
    if isinstance(x, Dual):
        return Dual(np.sin(x.v), np.cos(x.v)[..., None] * x.t)
    return np.sin(x)


def _hypot(a, b):
    if isinstance(a, Jet) or isinstance(b, Jet):
        like = a if isinstance(a, Jet) else b
        a, b = like._lift(a), like._lift(b)
        return (a * a + b * b).sqrt()
    if isinstance(a, Dual) or isinstance(b, Dual):
        like = a if isinstance(a, Dual) else b
        a, b = Dual.lift(a, like), Dual.lift(b, like)
        h = np.hypot(a.v, b.v)
        with np.errstate(invalid="ignore", divide="ignore"):
            t = (a.v / h)[..., None] * a.t + (b.v / h)[..., None] * b.t
        return Dual(h, t)
    return np.hypot(a, b)


def _arctan2(y, x):
    if isinstance(y, Jet) or isinstance(x, Jet):
        like = y if isinstance(y, Jet) else x
        y, x = like._lift(y), like._lift(x)
        y0, x0 = y.c[:, 0], x.c[:, 0]
        # rotate by -phi0: u = tan(phi - phi0) has no constant term, atan(u) = u - u^3/3 + O(u^5)
        u = (y * x0 - x * y0) / (x * x0 + y * y0)
        u.c[:, 0] = 0.0
        return Jet.const(np.arctan2(y0, x0), like) + u - (u * u * u) * (1.0 / 3.0)
    if isinstance(y, Dual) or isinstance(x, Dual):
        like = y if isinstance(y, Dual) else x
        y, x = Dual.lift(y, like), Dual.lift(x, like)
        r2 = x.v * x.v + y.v * y.v
        with np.errstate(invalid="ignore", divide="ignore"):
            t = (x.v / r2)[..., None] * y.t - (y.v / r2)[..., None] * x.t
        return Dual(np.arctan2(y.v, x.v), t)
    return np.arctan2(y, x)


def _where_zero(a, eps):
    # jnp.where(a == 0, eps, a)  (aberrations.py:101-102)
    if isinstance(a, Jet):
        m = a.c[:, 0] == 0
        c = np.where(m[:, None], 0.0, a.c)
        c[:, 0] = np.where(m, eps, a.c[:, 0])
        return Jet(c, a.space)
    if isinstance(a, Dual):
        m = a.v == 0
        return Dual(np.where(m, eps, a.v), np.where(m[..., None], 0.0, a.t))
    return np.where(a == 0, eps, a)


# --------------------------------------------------------------------------
# Ray (ray.py:8-45)
# --------------------------------------------------------------------------

RAY_FIELDS = ("x", "y", "dx", "dy", "z", "pathlength", "_one")


@dataclass
class Ray:
    x: Any
    y: Any
    dx: Any
    dy: Any
    z: Any
    pathlength: Any
    _one: Any = 1.0

    def derive(self, **kw):  # ray.py:127-166
        return replace(self, **kw)

    @staticmethod
    def from_obj(obj) -> "Ray":
        """Build from any object with the seven Ray attributes (duck typing)."""
        def get(name):
            v = getattr(obj, name)
            if hasattr(v, "detach"):  # torch tensor
                v = v.detach().cpu().numpy()
            return v
        return Ray(*(get(f) for f in RAY_FIELDS))

    def as_arrays(self):
        """All seven fields as float64 arrays of one common length (``to_vector``)."""
        vals = [np.asarray(getattr(self, f), dtype=np.float64).reshape(-1) for f in RAY_FIELDS]
        n = max(v.size for v in vals)
        return Ray(*(np.broadcast_to(v, (n,)).copy() for v in vals))


# --------------------------------------------------------------------------
# propagation and components
# --------------------------------------------------------------------------


def propagate(ray: Ray, distance) -> Ray:
    """FreeSpaceParaxial.propagate  (propagator.py:52-72)."""
    return ray.derive(
        x=ray.x + ray.dx * distance,
        y=ray.y + ray.dy * distance,
        z=ray.z + distance,
        pathlength=ray.pathlength + distance,
    )


_KRIV_FIELDS = (
    "C10", "C12", "phi12", "C21", "phi21", "C23", "phi23", "C30", "C32", "phi32",
    "C34", "phi34", "C41", "phi41", "C43", "phi43", "C45", "phi45", "C50", "C52",
    "phi52", "C54", "phi54", "C56", "phi56",
)


def _ck(m, ph, ph0):  # aberrations.py:34-35
    return _cos(m * (ph - ph0))


def _sk(m, ph, ph0):  # aberrations.py:38-39
    return _sin(m * (ph - ph0))


def krivanek_coeff_brackets(phi, p):  # aberrations.py:42-48
    B2 = p.C10 + p.C12 * _ck(2, phi, p.phi12)
    B3 = p.C21 * _ck(1, phi, p.phi21) + p.C23 * _ck(3, phi, p.phi23)
    B4 = p.C30 + p.C32 * _ck(2, phi, p.phi32) + p.C34 * _ck(4, phi, p.phi34)
    B5 = (p.C41 * _ck(1, phi, p.phi41) + p.C43 * _ck(3, phi, p.phi43)
          + p.C45 * _ck(5, phi, p.phi45))
    B6 = (p.C50 + p.C52 * _ck(2, phi, p.phi52) + p.C54 * _ck(4, phi, p.phi54)
          + p.C56 * _ck(6, phi, p.phi56))
    return B2, B3, B4, B5, B6


def W_krivanek(alpha, phi, p):  # aberrations.py:51-60
    B2, B3, B4, B5, B6 = krivanek_coeff_brackets(phi, p)
    a = alpha
    a2 = a * a
    a3 = a2 * a
    a4 = a2 * a2
    a6 = a3 * a3
    return 0.5 * a2 * B2 + (a3 / 3.0) * B3 + 0.25 * a4 * B4 + 0.2 * a4 * a * B5 + (a6 / 6.0) * B6


def grad_W_krivanek(alpha_x, alpha_y, p):  # aberrations.py:63-108
    ax, ay = alpha_x, alpha_y
    alpha = _hypot(ax, ay)
    phi = _arctan2(ay, ax)
    B2, B3, B4, B5, B6 = krivanek_coeff_brackets(phi, p)
    a = alpha
    a2 = a * a
    a3 = a2 * a
    a4 = a2 * a2
    a5 = a4 * a
    a6 = a3 * a3
    dW_dalpha = a * B2 + a2 * B3 + a3 * B4 + a4 * B5 + a5 * B6
    dW_dphi = (0.5 * a2) * (-2.0 * p.C12 * _sk(2, phi, p.phi12))
    dW_dphi = dW_dphi + (a3 / 3.0) * (
        -1.0 * p.C21 * _sk(1, phi, p.phi21) - 3.0 * p.C23 * _sk(3, phi, p.phi23))
    dW_dphi = dW_dphi + (0.25 * a4) * (
        -2.0 * p.C32 * _sk(2, phi, p.phi32) - 4.0 * p.C34 * _sk(4, phi, p.phi34))
    dW_dphi = dW_dphi + (0.2 * a4 * a) * (
        -1.0 * p.C41 * _sk(1, phi, p.phi41) - 3.0 * p.C43 * _sk(3, phi, p.phi43)
        - 5.0 * p.C45 * _sk(5, phi, p.phi45))
    dW_dphi = dW_dphi + (a6 / 6.0) * (
        -2.0 * p.C52 * _sk(2, phi, p.phi52) - 4.0 * p.C54 * _sk(4, phi, p.phi54)
        - 6.0 * p.C56 * _sk(6, phi, p.phi56))
    eps = 1e-30
    a_safe = _where_zero(a, eps)
    inv_a = 1.0 / a_safe
    inv_a2 = inv_a * inv_a
    dWx = dW_dalpha * (ax * inv_a) + dW_dphi * (-ay * inv_a2)
    dWy = dW_dalpha * (ay * inv_a) + dW_dphi * (ax * inv_a2)
    return dWx, dWy


def component_z(comp):
    """``component.z`` (ThickLens: property returning z_po, components.py:449-452)."""
    if type(comp).__name__ == "ThickLens":
        return comp.z_po
    return comp.z


def apply_component(comp, ray: Ray) -> Ray:
    """``component(ray)`` for every component class of components.py / source.py."""
    name = type(comp).__name__
    if name in ("Plane", "Detector", "ScanGrid", "PointSource", "ParallelBeam", "Source"):
        # components.py:133-134, 248-249, 405-406; source.py:21-34
        return ray
    if name == "Lens":  # components.py:161-174
        f = comp.focal_length
        x, y, dx, dy = ray.x, ray.y, ray.dx, ray.dy
        new_dx = -x / f + dx
        new_dy = -y / f + dy
        pathlength = ray.pathlength - (x ** 2 + y ** 2) / (2 * f)
        one = ray._one * 1.0
        return Ray(x=x, y=y, dx=new_dx, dy=new_dy, _one=one, pathlength=pathlength, z=ray.z)
    if name == "ThickLens":  # components.py:431-447
        f = comp.focal_length
        x, y, dx, dy = ray.x, ray.y, ray.dx, ray.dy
        new_dx = -x / f + dx
        new_dy = -y / f + dy
        pathlength = ray.pathlength - (x ** 2 + y ** 2) / (2 * f)
        new_z = ray.z - (comp.z_po - comp.z_pi)
        one = ray._one * 1.0
        return Ray(x=x, y=y, dx=new_dx, dy=new_dy, _one=one, pathlength=pathlength, z=new_z)
    if name == "AberratedLensKrivanek":  # components.py:192-215
        f = comp.focal_length
        x, y, dx, dy = ray.x, ray.y, ray.dx, ray.dy
        coeffs = comp.coeffs
        ideal_dx = -x / f + dx
        ideal_dy = -y / f + dy
        alpha = _hypot(ideal_dx, ideal_dy)
        phi = _arctan2(ideal_dy, ideal_dx)
        dWx, dWy = grad_W_krivanek(ideal_dx, ideal_dy, coeffs)
        dux, duy = -dWx / f, -dWy / f
        aber_dx = ideal_dx + dux
        aber_dy = ideal_dy + duy
        pathlength = (ray.pathlength - (x ** 2 + y ** 2) / (2 * f)
                      + W_krivanek(alpha, phi, coeffs) / f)
        one = ray._one * 1.0
        return Ray(x=x, y=y, dx=aber_dx, dy=aber_dy, _one=one, pathlength=pathlength, z=ray.z)
    if name == "Deflector":  # components.py:476-482
        x, y, dx, dy = ray.x, ray.y, ray.dx, ray.dy
        return ray.derive(
            dx=dx + comp.def_x * ray._one,
            dy=dy + comp.def_y * ray._one,
            pathlength=ray.pathlength + dx * x + dy * y,
        )
    if name == "Biprism":  # components.py:553-559 (offset, rotation, side unused)
        x, y, dx, dy = ray.x, ray.y, ray.dx, ray.dy
        return ray.derive(
            dx=dx + comp.def_x * ray._one * _sign(ray.x),
            dy=dy,
            pathlength=ray.pathlength + dx * x + dy * y,
        )
    if name == "Scanner":  # components.py:279-285
        return ray.derive(
            x=ray.x + comp.scan_pos_x * ray._one,
            y=ray.y + comp.scan_pos_y * ray._one,
            dx=ray.dx + comp.scan_tilt_x * ray._one,
            dy=ray.dy + comp.scan_tilt_y * ray._one,
        )
    if name == "Descanner":  # components.py:343-372
        de = comp.descan_error
        sp_x, sp_y = comp.scan_pos_x, comp.scan_pos_y
        st_x, st_y = comp.scan_tilt_x, comp.scan_tilt_y
        return ray.derive(
            x=ray.x + (sp_x * de.pxo_pxi + sp_y * de.pxo_pyi + de.offpxi - sp_x) * ray._one,
            y=ray.y + (sp_x * de.pyo_pxi + sp_y * de.pyo_pyi + de.offpyi - sp_y) * ray._one,
            dx=ray.dx + (sp_x * de.sxo_pxi + sp_y * de.sxo_pyi + de.offsxi - st_x) * ray._one,
            dy=ray.dy + (sp_x * de.syo_pxi + sp_y * de.syo_pyi + de.offsyi - st_y) * ray._one,
        )
    if name == "Rotator":  # components.py:503-523
        angle = comp.angle * (np.pi / 180.0)  # jnp.deg2rad
        c, s = _cos(angle), _sin(angle)
        return Ray(
            x=ray.x * c - ray.y * s,
            y=ray.x * s + ray.y * c,
            dx=ray.dx * c - ray.dy * s,
            dy=ray.dx * s + ray.dy * c,
            _one=ray._one,
            pathlength=ray.pathlength,
            z=ray.z,
        )
    raise TypeError(f"oracle: unknown component {name}")


_SOURCE_OR_COMPONENT = (
    "Plane", "Detector", "ScanGrid", "PointSource", "ParallelBeam", "Source", "Lens",
    "ThickLens", "AberratedLensKrivanek", "Deflector", "Biprism", "Scanner", "Descanner",
    "Rotator",
)


def run_iter(ray: Ray, components: Sequence[Any]):
    """run.py:47-82 -- a propagation step is ALWAYS inserted (also for distance 0)."""
    for comp in components:
        distance = component_z(comp) - ray.z  # run.py:77
        ray = propagate(ray, distance)  # run.py:78-79
        yield ("Propagator", distance), ray
        ray = apply_component(comp, ray)  # run.py:81
        yield comp, ray


def run_to_end(ray, components: Sequence[Any]) -> Ray:
    """run.py:85-116."""
    if not isinstance(ray, Ray):
        ray = Ray.from_obj(ray)
    for _, ray in run_iter(ray, components):
        pass
    return ray


def _seed_duals(ray) -> Ray:
    if not isinstance(ray, Ray):
        ray = Ray.from_obj(ray)
    r = ray.as_arrays()
    n = r.x.shape[0]
    out = []
    for i, f in enumerate(RAY_FIELDS):
        t = np.zeros((n, 7))
        t[:, i] = 1.0
        out.append(Dual(getattr(r, f), t))
    return Ray(*out)


def jacobian_run_to_end(ray, components: Sequence[Any]):
    """``vmap(jacobian(run_to_end), in_axes=(0, None))(rays, model)`` (gaussian.py:234-236).

    Returns ``(out_ray_values (Ray of (N,) arrays), J (N,7,7))`` with
    ``J[n, i, j] = d out_i / d in_j`` in the field order ``RAY_FIELDS``.
    """
    if not isinstance(ray, Ray):
        ray = Ray.from_obj(ray)
    d = _seed_duals(ray)
    out = run_to_end(d, components)
    n = d.x.v.shape[0]
    J = np.zeros((n, 7, 7))
    vals = []
    for i, f in enumerate(RAY_FIELDS):
        o = getattr(out, f)
        if not isinstance(o, Dual):
            o = Dual(np.broadcast_to(np.asarray(o, dtype=np.float64), (n,)), np.zeros((n, 7)))
        J[:, i, :] = o.t
        vals.append(np.array(o.v))
    return Ray(*vals), J


def calculate_derivatives(ray, components: Sequence[Any], order: int):
    """``calculate_derivatives(ray, model, order)`` (run.py:119-147): the k-th entry is what ``order``
    nested ``jax.jacfwd(run_to_end, argnums=0)`` calls produce, as a dense array
    ``D_k[n, f, a_1, ..., a_k] = d^k out_f / d in_a1 ... d in_ak`` in ``RAY_FIELDS`` order
    (the reference returns the same numbers as nested Ray pytrees: ``derivs[1].x.dx.dy``)."""
    if not isinstance(ray, Ray):
        ray = Ray.from_obj(ray)
    r = ray.as_arrays()
    n = r.x.shape[0]
    space = JetSpace(7, order)
    seeded = Ray(*(Jet.variable(getattr(r, f), i, space) for i, f in enumerate(RAY_FIELDS)))
    out = run_to_end(seeded, components)
    derivs = []
    for k in range(1, order + 1):
        Dk = np.zeros((n, 7) + (7,) * k)
        for i, f in enumerate(RAY_FIELDS):
            o = getattr(out, f)
            if isinstance(o, Jet):
                Dk[:, i] = o.derivative_tensor(k)
        derivs.append(Dk)
    return derivs


def run_with_grads(ray, components: Sequence[Any], directions):
    """``run_with_grads`` (run.py:182-267) restated with forward-mode duals.

    directions: list of ``("ray", field)`` or ``(component_index, (attr, ...))`` -- one tangent
    lane each.  Returns ``(out Ray of (N,) arrays, J (N, 7, K))``, J[n, i, k] = d out_i / d dir_k.
    """
    import dataclasses as _dc
    if not isinstance(ray, Ray):
        ray = Ray.from_obj(ray)
    r = ray.as_arrays()
    n = r.x.shape[0]
    K = len(directions)

    def seeded(value, k):
        t = np.zeros((n, K))
        if k is not None:
            t[:, k] = 1.0
        return Dual(np.broadcast_to(np.asarray(value, dtype=np.float64), (n,)).copy(), t)

    ray_k = {f: None for f in RAY_FIELDS}
    comps = list(components)
    for k, d in enumerate(directions):
        if d[0] == "ray":
            ray_k[d[1]] = k
        else:
            ci, path = d
            c = comps[ci]
            if len(path) == 1:
                c = _dc.replace(c, **{path[0]: seeded(getattr(c, path[0]), k)})
            else:  # ("descan_error", name) -- a NamedTuple -- or ("coeffs", name) -- a dataclass
                inner = getattr(c, path[0])
                new = {path[1]: seeded(getattr(inner, path[1]), k)}
                inner = inner._replace(**new) if hasattr(inner, "_replace") else _dc.replace(inner, **new)
                c = _dc.replace(c, **{path[0]: inner})
            comps[ci] = c
    d_ray = Ray(*(seeded(getattr(r, f), ray_k[f]) for f in RAY_FIELDS))
    out = run_to_end(d_ray, comps)
    J = np.zeros((n, 7, K))
    vals = []
    for i, f in enumerate(RAY_FIELDS):
        o = getattr(out, f)
        if not isinstance(o, Dual):
            o = Dual(np.broadcast_to(np.asarray(o, dtype=np.float64), (n,)), np.zeros((n, K)))
        J[:, i, :] = o.t
        vals.append(np.array(o.v))
    return Ray(*vals), J


_PICK = (0, 1, 2, 3, 6)


def custom_jacobian_matrix(J7: np.ndarray) -> np.ndarray:
    """utils.py:7-43 -- the ``[x, y, dx, dy, _one]`` 5x5 block of the 7x7 Jacobian."""
    J7 = np.asarray(J7)
    return J7[..., _PICK, :][..., :, _PICK]


def abcd_run_to_end(ray, components):
    """(out_ray, ABCD (N,5,5)) -- gaussian.py:234-239."""
    out, J = jacobian_run_to_end(ray, components)
    return out, custom_jacobian_matrix(J)


def solve_model(ray, components) -> np.ndarray:
    """run.py:150-179 -- per-step 5x5 Jacobians, shape (2*n_comp, 5, 5) for a scalar ray."""
    if not isinstance(ray, Ray):
        ray = Ray.from_obj(ray)
    mats = []
    for comp in components:
        d = _seed_duals(ray)
        distance = component_z(comp) - d.z
        o = propagate(d, distance)
        mats.append(_ray_jac(o))
        ray = Ray(*(getattr(o, f).v if isinstance(getattr(o, f), Dual) else getattr(o, f)
                    for f in RAY_FIELDS))
        d = _seed_duals(ray)
        o = apply_component(comp, d)
        mats.append(_ray_jac(o))
        ray = Ray(*(getattr(o, f).v if isinstance(getattr(o, f), Dual) else getattr(o, f)
                    for f in RAY_FIELDS))
    out = np.stack(mats, axis=0)  # (steps, N, 5, 5)
    if out.shape[1] == 1:
        out = out[:, 0]
    return out


def _ray_jac(o: Ray) -> np.ndarray:
    n = None
    for f in RAY_FIELDS:
        v = getattr(o, f)
        if isinstance(v, Dual):
            n = v.v.shape[0]
            break
    J = np.zeros((n, 7, 7))
    for i, f in enumerate(RAY_FIELDS):
        v = getattr(o, f)
        if isinstance(v, Dual):
            J[:, i, :] = v.t
    return custom_jacobian_matrix(J)


# --------------------------------------------------------------------------
# grid / coordinate transforms
# --------------------------------------------------------------------------


def _rotate(radians):  # coordinate_transforms.py:17-30
    c, s = np.cos(radians), np.sin(radians)
    return np.array([(c, s, 0.0), (-s, c, 0.0), (0.0, 0.0, 1.0)])


def _scale(ps):  # coordinate_transforms.py:33-36
    return np.array([(ps[0], 0.0, 0.0), (0.0, ps[1], 0.0), (0.0, 0.0, 1.0)])


def _shift(c):  # coordinate_transforms.py:39-42
    return np.array([(1.0, 0.0, c[0]), (0.0, 1.0, c[1]), (0.0, 0.0, 1.0)])


def _flip_y():  # coordinate_transforms.py:45-47
    return np.array([(-1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0)])


def pixels_to_metres_transform(centre, pixel_size, shape, flip_y=False, rotation=0.0):
    """coordinate_transforms.py:50-101 (composition order of lines 92-99)."""
    flip_transform = _flip_y() if flip_y else np.eye(3)
    shape = np.array(shape)
    return (
        flip_transform
        @ _flip_y()
        @ _rotate(np.pi / 180 * rotation)
        @ _shift(centre)
        @ _scale(pixel_size)
        @ _shift(-(shape - 1) / 2.0)
    )


def grid_pixels_to_metres_mat(grid):  # grid.py:37-48
    return pixels_to_metres_transform(grid.centre, grid.pixel_size, grid.shape,
                                      grid.flip_y, grid.rotation)


def grid_metres_to_pixels_mat(grid):  # grid.py:50-63
    return np.linalg.inv(grid_pixels_to_metres_mat(grid))


def apply_transformation(y, x, T):
    """coordinate_transforms.py:104-140: ``T @ [y, x, 1]``.

    The three-term dot is evaluated left to right without FMA contraction,
    ``(T[i,0]*y + T[i,1]*x) + T[i,2]*1`` -- this fixed order is the oracle's
    definition of the bit pattern (the reference leaves it to XLA's dot).
    """
    y = np.asarray(y, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    yt = (T[0, 0] * y + T[0, 1] * x) + T[0, 2]
    xt = (T[1, 0] * y + T[1, 1] * x) + T[1, 2]
    return yt, xt


def grid_pixels_to_metres(grid, pixels):  # grid.py:155-182
    py, px = pixels
    my, mx = apply_transformation(py, px, grid_pixels_to_metres_mat(grid))
    return mx, my


def round_to_int32(v):
    """``jnp.round(v).astype(int32)`` (grid.py:147-149): half-to-even, saturating,
    NaN -> 0 (XLA convert semantics; also what ``cvt.rni.s32.f64`` does)."""
    r = np.rint(np.asarray(v, dtype=np.float64))
    r = np.where(np.isnan(r), 0.0, r)
    r = np.clip(r, -2147483648.0, 2147483647.0)
    return r.astype(np.int32)


def grid_metres_to_pixels(grid, coords, cast=True):  # grid.py:120-153
    cx, cy = coords
    py, px = apply_transformation(cy, cx, grid_metres_to_pixels_mat(grid))
    if cast:
        py, px = round_to_int32(py), round_to_int32(px)
    return py, px


def grid_coords(grid) -> np.ndarray:
    """grid.py:65-100 -- (H*W, 2) of (x_m, y_m), row-major."""
    H, W = grid.shape
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    cx, cy = grid_pixels_to_metres(grid, (yy.ravel(), xx.ravel()))
    return np.stack((cx, cy), axis=-1).reshape(-1, 2)


def inplace_sum(px_y, px_x, mask, frame, buffer):
    """utils.py:83-114 -- bounds-checked scatter add."""
    h, w = buffer.shape
    ok = mask & (px_y >= 0) & (px_y < h) & (px_x >= 0) & (px_x < w)
    np.add.at(buffer, (px_y[ok], px_x[ok]), frame[ok].astype(buffer.dtype))


def grid_into_image(grid, ray, acc=None):  # grid.py:231-280
    yy, xx = grid_metres_to_pixels(grid, (np.atleast_1d(ray.x), np.atleast_1d(ray.y)), cast=True)
    if acc is None:
        acc = np.zeros(tuple(grid.shape), dtype=int)
    inplace_sum(yy, xx, np.ones(yy.shape, bool), np.ones(xx.shape, np.float32), acc)
    return acc


# --------------------------------------------------------------------------
# Gaussian beamlets
# --------------------------------------------------------------------------


def _solve2(Amat, Bmat):
    """Batched 2x2 ``jnp.linalg.solve(A, B)`` restated as LU with partial pivoting
    (LAPACK gesv order).  Singular pivots produce inf/nan rather than raising,
    like JAX.  Amat, Bmat: (..., 2, 2)."""
    A = np.asarray(Amat)
    B = np.asarray(Bmat)
    dt = np.result_type(A.dtype, B.dtype, np.float64)
    A = A.astype(dt)
    B = np.broadcast_to(B.astype(dt), np.broadcast_shapes(A.shape, B.shape)).copy()
    A = np.broadcast_to(A, B.shape).copy()
    swap = np.abs(A[..., 1, 0]) > np.abs(A[..., 0, 0])
    A[swap] = A[swap][..., ::-1, :]
    B[swap] = B[swap][..., ::-1, :]
    with np.errstate(all="ignore"):
        l10 = A[..., 1, 0] / A[..., 0, 0]
        u11 = A[..., 1, 1] - l10 * A[..., 0, 1]
        y1 = B[..., 1, :] - l10[..., None] * B[..., 0, :]
        x1 = y1 / u11[..., None]
        x0 = (B[..., 0, :] - A[..., 0, 1][..., None] * x1) / A[..., 0, 0][..., None]
    return np.stack([x0, x1], axis=-2)


def _det2(A):
    """``jnp.linalg.det`` of (...,2,2) via the LU diagonal."""
    A = np.asarray(A)
    swap = np.abs(A[..., 1, 0]) > np.abs(A[..., 0, 0])
    a00 = np.where(swap, A[..., 1, 0], A[..., 0, 0])
    a01 = np.where(swap, A[..., 1, 1], A[..., 0, 1])
    a10 = np.where(swap, A[..., 0, 0], A[..., 1, 0])
    a11 = np.where(swap, A[..., 0, 1], A[..., 1, 1])
    with np.errstate(all="ignore"):
        u11 = a11 - (a10 / a00) * a01
        d = a00 * u11
    return np.where(swap, -d, d)


def _nan_to_num0(a):  # jnp.nan_to_num(a, nan=0, posinf=0, neginf=0) (gaussian.py:284, 287)
    a = np.array(a)
    if np.iscomplexobj(a):
        re = np.where(np.isfinite(a.real), a.real, 0.0)
        im = np.where(np.isfinite(a.imag), a.imag, 0.0)
        return re + 1j * im
    return np.where(np.isfinite(a), a, 0.0)


def gaussian_q_inv(waist_xy, radii_of_curv, wavelength):
    """GaussianRay.q_inv  (gaussian.py:138-155)."""
    waist_xy = np.atleast_2d(np.asarray(waist_xy, dtype=np.float64))
    radii = np.atleast_2d(np.asarray(radii_of_curv, dtype=np.float64))
    wl = np.asarray(wavelength, dtype=np.float64)
    w_x, w_y = waist_xy.T
    R_x, R_y = radii.T
    with np.errstate(all="ignore"):
        inv_qx = np.where(np.isinf(R_x), 1j * wl / (np.pi * w_x ** 2),
                          -1.0 / R_x + 1j * wl / (np.pi * w_x ** 2))
        inv_qy = np.where(np.isinf(R_y), 1j * wl / (np.pi * w_y ** 2),
                          -1.0 / R_y + 1j * wl / (np.pi * w_y ** 2))
    return inv_qx, inv_qy


def gaussian_Q_inv(waist_xy, radii_of_curv, wavelength, theta):
    """GaussianRay.Q_inv  (gaussian.py:157-177): einsum("nij,njk,npk->nip", R, D, R)."""
    inv_qx, inv_qy = gaussian_q_inv(waist_xy, radii_of_curv, wavelength)
    theta = np.atleast_1d(np.asarray(theta, dtype=np.float64))
    n = inv_qx.shape[0]
    D = np.zeros((n, 2, 2), dtype=np.complex128)
    D[:, 0, 0] = inv_qx
    D[:, 1, 1] = inv_qy
    c, s = np.cos(theta), np.sin(theta)
    R = np.zeros((n, 2, 2))
    R[:, 0, 0] = c
    R[:, 0, 1] = -s
    R[:, 1, 0] = s
    R[:, 1, 1] = c
    return np.einsum("nij,njk,npk->nip", R, D, R)


def Qinv_ABCD(Qinv, A, B, C, D):
    """gaussian.py:92-96: solve(A + B Qinv, C + D Qinv)."""
    lhs = A + B @ Qinv
    rhs = C + D @ Qinv
    return _solve2(lhs, rhs)


def _beam_field_batch(amp, phase_offset, Q1_inv, Q2_inv, r1m, theta1m, A, B, e, f, k, r2):
    """``_beam_field`` (gaussian.py:276-316) for a batch of beamlets at once.

    Shapes: amp, phase_offset, k (nb,); Q1_inv, Q2_inv, A, B (nb,2,2); r1m,
    theta1m, e, f (nb,2); r2 (npix,2).  Returns (nb, npix) complex128.
    """
    I = np.eye(2)
    B_inv = _nan_to_num0(_solve2(B, I))                       # :283-284
    Q1 = _nan_to_num0(_solve2(Q1_inv, I.astype(np.complex128)))  # :286-287
    r2s = r2[None, :, :] - e[:, None, :]                      # :289
    r2m = np.einsum("nij,nj->ni", A, r1m) + np.einsum("nij,nj->ni", B, theta1m)  # :291
    denom = A + np.einsum("nij,njk->nik", B, Q1_inv)          # :294
    with np.errstate(all="ignore"):
        pref = amp / np.sqrt(_det2(denom))                    # :295
    ABinv = np.einsum("nij,njk->nik", A, B_inv)               # :298
    phi1 = (np.einsum("ni,nij,nj->n", r1m, ABinv, r1m)[:, None]
            - 2 * np.einsum("ni,nij,npj->np", r1m, B_inv, r2s))          # :299-301
    AQ1 = np.einsum("nij,njk->nik", A, Q1)                    # :304
    B_over_AQ1B = _solve2(np.einsum("nij,njk->nik", B, AQ1 + B), I.astype(np.complex128))  # :305
    Q1B_over_AQ = np.einsum("nij,njk->nik", Q1, B_over_AQ1B)  # :306
    phi2 = (np.einsum("ni,nij,nj->n", r2m, Q1B_over_AQ, r2m)[:, None]
            - 2 * np.einsum("ni,nij,npj->np", r2m, Q1B_over_AQ, r2s))    # :307-309
    Q2t = np.einsum("npi,nij,npj->np", r2s, Q2_inv, r2s)      # :311
    f_offset = 2 * np.einsum("npj,nj->np", r2s, f)            # :314
    phase = (k / 2)[:, None] * (Q2t + phi1 - phi2 + f_offset)  # :315
    with np.errstate(all="ignore"):
        return pref[:, None] * np.exp(1j * (phase + phase_offset[:, None]))  # :316


def propagate_misaligned_gaussian(amp, phase_offset, Q1_inv, A, B, C, D, e, f, r1m, theta1m,
                                  k, r2, batch_size=128, pix_block=1 << 16):
    """``propagate_misaligned_gaussian_jax_scan`` (gaussian.py:319-337) with the
    ``map_reduce`` order (gaussian.py:340-369): beamlets are added one by one in
    natural order into a complex128 accumulator."""
    amp = np.asarray(amp, dtype=np.float64)
    nb = amp.shape[0]
    r2 = np.asarray(r2, dtype=np.float64)
    npix = r2.shape[0]
    Q1_inv = np.asarray(Q1_inv, dtype=np.complex128)
    Q2_inv = Qinv_ABCD(Q1_inv, A, B, C, D)                    # :323
    out = np.zeros((npix,), dtype=np.complex128)
    for p0 in range(0, npix, pix_block):
        p1 = min(npix, p0 + pix_block)
        acc = np.zeros((p1 - p0,), dtype=np.complex128)
        for b0 in range(0, nb, batch_size):
            s = slice(b0, min(nb, b0 + batch_size))
            fld = _beam_field_batch(amp[s], phase_offset[s], Q1_inv[s], Q2_inv[s], r1m[s],
                                    theta1m[s], A[s], B[s], e[s], f[s], k[s], r2[p0:p1])
            for row in fld:           # sequential add, gaussian.py:348-355
                acc += row
        out[p0:p1] = acc
    return out


def _input_beam_field_batch(a, p, q1, r1m, t1m, k, r1):
    """``_input_beam_field`` (gaussian.py:402-407) for a batch; returns (nb, npix)."""
    d = r1[None, :, :] - r1m[:, None, :]
    quad = np.einsum("npi,nij,npj->np", d, q1, d)
    tilt = 2 * np.einsum("npj,nj->np", d, t1m)
    phase = (k / 2)[:, None] * (quad + tilt)
    with np.errstate(all="ignore"):
        return a[:, None] * np.exp(1j * phase) * np.exp(1j * p)[:, None]


def evaluate_misaligned_input_gaussian(amp, phase_offset, Q1_inv, r1m, theta1m, k, r1,
                                       batch_size=128, pix_block=1 << 16):
    """gaussian.py:410-427."""
    amp = np.asarray(amp, dtype=np.float64)
    nb = amp.shape[0]
    npix = r1.shape[0]
    out = np.zeros((npix,), dtype=np.complex128)
    for p0 in range(0, npix, pix_block):
        p1 = min(npix, p0 + pix_block)
        acc = np.zeros((p1 - p0,), dtype=np.complex128)
        for b0 in range(0, nb, batch_size):
            s = slice(b0, min(nb, b0 + batch_size))
            fld = _input_beam_field_batch(amp[s], phase_offset[s], Q1_inv[s], r1m[s],
                                          theta1m[s], k[s], r1[p0:p1])
            for row in fld:
                acc += row
        out[p0:p1] = acc
    return out


def _gr_arrays(g):
    """Pull numpy arrays out of a GaussianRay-like object (``to_vector``, gaussian.py:229)."""
    def arr(name):
        v = getattr(g, name)
        if hasattr(v, "detach"):
            v = v.detach().cpu().numpy()
        return np.atleast_1d(np.asarray(v, dtype=np.float64))
    d = {f: arr(f) for f in RAY_FIELDS}
    d["amplitude"] = arr("amplitude")
    d["wavelength"] = arr("wavelength")
    d["theta"] = arr("theta")
    w = getattr(g, "waist_xy")
    r = getattr(g, "radii_of_curv")
    if hasattr(w, "detach"):
        w = w.detach().cpu().numpy()
    if hasattr(r, "detach"):
        r = r.detach().cpu().numpy()
    d["waist_xy"] = np.atleast_2d(np.asarray(w, dtype=np.float64))
    d["radii_of_curv"] = np.atleast_2d(np.asarray(r, dtype=np.float64))
    return d


def make_gaussian_image(gaussian_rays, model, batch_size=128):
    """gaussian.py:225-273."""
    g = _gr_arrays(gaussian_rays)
    grid = model[-1]
    central = Ray(*(g[f] for f in RAY_FIELDS))
    _, J = jacobian_run_to_end(central, model)                # :234-236
    ABCDs = custom_jacobian_matrix(J)                          # :238-239
    n = ABCDs.shape[0]
    Q1_invs = gaussian_Q_inv(g["waist_xy"], g["radii_of_curv"], g["wavelength"], g["theta"])
    As, Bs = ABCDs[:, 0:2, 0:2], ABCDs[:, 0:2, 2:4]           # :244-245
    Cs, Ds = ABCDs[:, 2:4, 0:2], ABCDs[:, 2:4, 2:4]           # :246-247
    es, fs = ABCDs[:, 0:2, 4], ABCDs[:, 2:4, 4]               # :248-249
    r2 = grid_coords(grid)                                     # :250
    ca = central.as_arrays()
    r1ms = np.stack([np.broadcast_to(ca.x, (n,)), np.broadcast_to(ca.y, (n,))], axis=-1)
    theta1ms = np.stack([np.broadcast_to(ca.dx, (n,)), np.broadcast_to(ca.dy, (n,))], axis=-1)
    k = 2 * np.pi / np.broadcast_to(g["wavelength"], (n,))     # :253-254
    phase_offset = k * np.broadcast_to(g["pathlength"], (n,))  # :255
    amp = np.broadcast_to(g["amplitude"], (n,))
    out = propagate_misaligned_gaussian(amp, phase_offset, Q1_invs, As, Bs, Cs, Ds, es, fs,
                                        r1ms, theta1ms, k, r2, batch_size=batch_size)
    return out.reshape(tuple(grid.shape))


def evaluate_gaussian_input_image(gaussian_rays, grid, batch_size=128):
    """gaussian.py:372-399."""
    g = _gr_arrays(gaussian_rays)
    n = max(g["amplitude"].shape[0], g["x"].shape[0])
    Q1_invs = gaussian_Q_inv(g["waist_xy"], g["radii_of_curv"], g["wavelength"], g["theta"])
    Q1_invs = np.broadcast_to(Q1_invs, (n, 2, 2))
    r1 = grid_coords(grid)
    r1ms = np.stack([np.broadcast_to(g["x"], (n,)), np.broadcast_to(g["y"], (n,))], axis=-1)
    theta1ms = np.stack([np.broadcast_to(g["dx"], (n,)), np.broadcast_to(g["dy"], (n,))], axis=-1)
    wl = np.broadcast_to(g["wavelength"], (n,))                # :385
    k = 2 * np.pi / wl
    phase_offset = k * np.broadcast_to(g["pathlength"], (n,))
    amp = np.broadcast_to(g["amplitude"], (n,))
    out = evaluate_misaligned_input_gaussian(amp, phase_offset, Q1_invs, r1ms, theta1ms, k, r1,
                                             batch_size=batch_size)
    return out.reshape(tuple(grid.shape))


# closed-form helpers used by the reference tests as analytic anchors
def zR(w0, wavelength):  # gaussian.py:17-18
    return (np.pi * w0 ** 2) / wavelength


def w_z(w0, z, z_r):  # gaussian.py:13-14
    return w0 * np.sqrt(1 + (z / z_r) ** 2)


def fibonacci_spiral(nb_samples: int, radius: float, alpha=2):
    """utils.py:297-325 (numpy in the reference too) -- used to build the
    BASELINE synthetic beamlet layouts."""
    ga = np.pi * (3.0 - np.sqrt(5.0))
    np_boundary = np.round(alpha * np.sqrt(nb_samples))
    ii = np.arange(nb_samples)
    with np.errstate(invalid="ignore"):
        rr = np.where(
            ii > nb_samples - (np_boundary + 1),
            radius,
            radius * np.sqrt((ii + 0.5) / (nb_samples - 0.5 * (np_boundary + 1))),
        )
    rr[0] = 0.0
    phi = ii * ga
    return rr * np.cos(phi), rr * np.sin(phi)


# --------------------------------------------------------------------------
# transfer.py
# --------------------------------------------------------------------------


def accumulate_matrices_cumulative(matrices):
    """transfer.py:150-183 -- starts from the LAST matrix and left-multiplies backwards."""
    total = matrices[-1]
    out = [total]
    for tm in reversed(list(matrices[:-1])):
        total = tm @ total
        out.append(total)
    return np.stack(out, axis=0)


def transfer_rays(ray_coords, transfer_matrices):
    """transfer.py:6-54: einsum("mij,nj->nmi", cumulative, rays)."""
    cum = accumulate_matrices_cumulative(np.asarray(transfer_matrices, dtype=np.float64))
    return np.einsum("mij,nj->nmi", cum, np.asarray(ray_coords, dtype=np.float64))


# --------------------------------------------------------------------------
# 4D-STEM shadow-image backprojection, composed from the reference's building blocks
# (Scanner / Descanner components.py:252-372, Grid maps grid.py:120-182,
#  transfer_rays_pt_src transfer.py:57-123, inplace_sum utils.py:83-114).
# The composite itself is not in the reference: parity is per building block.
# --------------------------------------------------------------------------


def stem4d_pixel_indices(model_fn, scan_grid, detector, source_xy=(0.0, 0.0), out_grid=None, unrounded=False):
    """(Sy*Sx, Dy*Dx, 2) int32 sample-grid (py, px) of every (scan position, detector pixel) ray
    (``unrounded``: the fp64 pixel coordinates before ``round``, for tie statistics in tests)."""
    out_grid = scan_grid if out_grid is None else out_grid
    r0 = np.array([float(source_xy[0]), float(source_xy[1])])

    def mats(spx, spy):
        model = list(model_fn(spx, spy))
        idx = next(i for i, c in enumerate(model) if c is scan_grid)
        ray = Ray(x=r0[0], y=r0[1], dx=0.0, dy=0.0, z=float(component_z(model[0])), pathlength=0.0)
        return abcd_run_to_end(ray, model)[1][0], abcd_run_to_end(ray, model[:idx + 1])[1][0]

    d00, s00 = mats(0.0, 0.0)
    d10, s10 = mats(1.0, 0.0)
    d01, s01 = mats(0.0, 1.0)
    Adet, Bdet = d00[0:2, 0:2], d00[0:2, 2:4]
    Asamp, Bsamp = s00[0:2, 0:2], s00[0:2, 2:4]
    Binv = np.linalg.inv(Bdet)
    cdet, csamp = Adet @ r0, Asamp @ r0
    Sy, Sx = scan_grid.shape
    Dy, Dx = detector.shape
    sy, sx = np.meshgrid(np.arange(Sy), np.arange(Sx), indexing="ij")
    spy, spx = apply_transformation(sy.ravel(), sx.ravel(), grid_pixels_to_metres_mat(scan_grid))
    dy, dx = np.meshgrid(np.arange(Dy), np.arange(Dx), indexing="ij")
    yd, xd = apply_transformation(dy.ravel(), dx.ravel(), grid_pixels_to_metres_mat(detector))
    spx, spy = spx[:, None], spy[:, None]
    edx = (d00[0, 4] + spx * (d10[0, 4] - d00[0, 4])) + spy * (d01[0, 4] - d00[0, 4])
    edy = (d00[1, 4] + spx * (d10[1, 4] - d00[1, 4])) + spy * (d01[1, 4] - d00[1, 4])
    esx = (s00[0, 4] + spx * (s10[0, 4] - s00[0, 4])) + spy * (s01[0, 4] - s00[0, 4])
    esy = (s00[1, 4] + spx * (s10[1, 4] - s00[1, 4])) + spy * (s01[1, 4] - s00[1, 4])
    rx = (xd[None, :] - cdet[0]) - edx
    ry = (yd[None, :] - cdet[1]) - edy
    tx = Binv[0, 0] * rx + Binv[0, 1] * ry
    ty = Binv[1, 0] * rx + Binv[1, 1] * ry
    xs = (csamp[0] + (Bsamp[0, 0] * tx + Bsamp[0, 1] * ty)) + esx
    ys = (csamp[1] + (Bsamp[1, 0] * tx + Bsamp[1, 1] * ty)) + esy
    py, px = apply_transformation(ys, xs, grid_metres_to_pixels_mat(out_grid))
    if unrounded:
        return np.stack([py, px], axis=-1)
    return np.stack([round_to_int32(py), round_to_int32(px)], axis=-1)


def stem4d_backproject(data4d, model_fn, scan_grid, detector, source_xy=(0.0, 0.0), out_grid=None):
    """Shadow-image backprojection: out[py, px] += data (bounds-checked, utils.py:83-114)."""
    out_grid = scan_grid if out_grid is None else out_grid
    idx = stem4d_pixel_indices(model_fn, scan_grid, detector, source_xy, out_grid)
    Oy, Ox = out_grid.shape
    py, px = idx[..., 0].ravel(), idx[..., 1].ravel()
    vals = np.asarray(data4d, dtype=np.float64).ravel()
    ok = (py >= 0) & (py < Oy) & (px >= 0) & (px < Ox)
    out = np.zeros((Oy, Ox))
    np.add.at(out, (py[ok], px[ok]), vals[ok])
    return out


# ----------------------------------------------------------------------------------------------------
# Callers either side of the path: restated for the parity tests of their device versions


def multi_cumsum_inplace(values, partitions, start):
    """utils.py:46-80, statement for statement (plain loop: small inputs only)."""
    part_idx = 0
    current_part_len = partitions[part_idx]
    part_count = 0
    values[0] = start
    for v_idx in range(1, values.size):
        if current_part_len == part_count:
            part_count = 0
            part_idx += 1
            current_part_len = partitions[part_idx]
            values[v_idx] = start
        else:
            values[v_idx] += values[v_idx - 1]
            part_count += 1


def concentric_rings(num_points_approx, radius):
    """utils.py:117-175 -> (N, 2) array of (y, x)."""
    num_rings = max(1, int(np.floor((-1 + np.sqrt(1 + 4 * num_points_approx / np.pi)) / 2)))
    num_points_kth_ring = np.round(2 * np.pi * np.arange(1, num_rings + 1)).astype(int)
    num_rings = num_points_kth_ring.size
    points_per_unit = num_points_approx / num_points_kth_ring.sum()
    points_per_ring = np.round(num_points_kth_ring * points_per_unit).astype(int)
    radii = np.linspace(0, radius, num_rings + 1, endpoint=True)[1:]
    with np.errstate(divide="ignore"):
        div_angle = 2 * np.pi / points_per_ring
    params = np.stack((radii, div_angle), axis=0)
    repeats = points_per_ring.tolist()
    all_params = np.repeat(params, repeats, axis=-1)
    if all_params.shape[1]:
        multi_cumsum_inplace(all_params[1, :], points_per_ring, 0.0)
    all_radii = all_params[0, :]
    all_angles = all_params[1, :]
    return np.stack((all_radii * np.sin(all_angles), all_radii * np.cos(all_angles)), axis=-1)


def decompose_Q_inv(Q_inv, wavelength, eps=1e-12):
    """gaussian.py:35-89 -> (waist1, waist2, R1, R2, theta), larger waist first."""
    Q_inv = np.asarray(Q_inv, dtype=np.complex128)
    Sm = np.imag(Q_inv)
    Sm = 0.5 * (Sm + np.swapaxes(Sm, -1, -2))
    _, evecs = np.linalg.eigh(Sm)

    def _make_right_handed(E):
        detE = np.linalg.det(E)
        flip = np.where(detE < 0, -1.0, 1.0)[..., None]
        E = E.copy()
        E[..., :, 1] = E[..., :, 1] * flip
        return E

    evecs = _make_right_handed(evecs)
    Qd = np.swapaxes(evecs, -1, -2) @ Q_inv @ evecs
    qd = np.stack([Qd[..., 0, 0], Qd[..., 1, 1]], axis=-1)

    def _waists_from_diag(q):
        im = np.imag(q)
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.sqrt(np.where(np.abs(im) > eps, np.abs(wavelength / (np.pi * im)), np.inf))

    w = _waists_from_diag(qd)
    swap = w[..., 0] < w[..., 1]
    qd = np.where(swap[..., None], qd[..., ::-1], qd)
    evecs = np.where(swap[..., None, None], evecs[..., :, ::-1], evecs)
    evecs = _make_right_handed(evecs)
    w = _waists_from_diag(qd)
    re = np.real(qd)
    with np.errstate(divide="ignore", invalid="ignore"):
        R = np.where(np.abs(re) > eps, 1.0 / re, np.inf)
    theta = np.arctan2(evecs[..., 1, 0], evecs[..., 0, 0])
    return w[..., 0], w[..., 1], R[..., 0], R[..., 1], theta
