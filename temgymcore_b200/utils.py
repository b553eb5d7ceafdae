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


def custom_jacobian_matrix(ray_jac):
    """-> ``(..., 5, 5)`` over ``[x, y, dx, dy, _one]`` (utils.py:34-43)."""
    m = ray_jac.matrix if isinstance(ray_jac, RayJacobian) else ray_jac
    if m.shape[-1] == 5 and m.shape[-2] == 5:
        return m
    if m.shape[-1] != 7 or m.shape[-2] != 7:
        raise ValueError(f"expected a (...,7,7) or (...,5,5) Jacobian, got {tuple(m.shape)}")
    idx = list(_PICK)
    return m[..., idx, :][..., :, idx]


def fibonacci_spiral(nb_samples: int, radius: float, alpha=2):
    """Host-side beamlet-centre sampler (utils.py:297-325); input preparation only."""
    import numpy as np
    ga = np.pi * (3.0 - np.sqrt(5.0))
    np_boundary = np.round(alpha * np.sqrt(nb_samples))
    ii = np.arange(nb_samples)
    with np.errstate(invalid="ignore"):
        rr = np.where(ii > nb_samples - (np_boundary + 1), radius,
                      radius * np.sqrt((ii + 0.5) / (nb_samples - 0.5 * (np_boundary + 1))))
    rr[0] = 0.
    phi = ii * ga
    return rr * np.cos(phi), rr * np.sin(phi)
