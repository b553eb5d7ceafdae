"""``jax.ffi`` registration of the CUDA path (BASELINE.json north_star: "a thin jax.ffi custom-call
C-ABI layer, so the JAX tracing surface and jax.jit / jax.jacobian ... are unchanged").

UNTESTED IN THIS ENVIRONMENT: jax / jaxlib are not installed and cannot be (no network),
so importing this module raises ``ImportError`` here.  It is the reference-side glue a maintainer adds
once ``csrc/xla/build_xla_shim.sh`` has built ``libtemgym_b200_xla.so`` on a machine with jax; the
tested drop-in boundary is ``include/temgym_b200.h`` + the Python modules next to this one.

``run_to_end_soa(rays, model)`` takes / returns ``(7, N)`` fp64 arrays in Ray field order and carries a
``jax.custom_jvp`` rule whose tangents come from the SAME dual-number kernel (``TG_JAC_FULL7``), so
``jax.jacobian(run_to_end_soa)`` and ``jax.vmap`` work without JAX ever tracing the components
(reference run.py:85-116, README.md:227-234).
"""
from __future__ import annotations

import ctypes
import os

import jax                       # noqa: F401  (ImportError here == "not available", by design)
import jax.numpy as jnp

from . import _lib as L
from .run import compile_model

_HERE = os.path.dirname(os.path.abspath(__file__))
_shim = ctypes.CDLL(os.path.join(_HERE, "libtemgym_b200_xla.so"))
for _name, _sym in (("tg_trace", "TgTrace"), ("tg_field_sum", "TgFieldSum"),
                    ("tg_beamlet_coeffs", "TgBeamletCoeffs")):
    jax.ffi.register_ffi_target(_name, jax.ffi.pycapsule(getattr(_shim, _sym)), platform="CUDA")


def _model_bytes(model) -> bytes:
    return bytes(compile_model(model))                     # tg_model is plain data (include/temgym_b200.h)


def _trace(rays, model_bytes: bytes, jac_layout: int):
    n = rays.shape[1]
    jdim = {L.TG_JAC_NONE: 0, L.TG_JAC_ABCD5: 5, L.TG_JAC_FULL7: 7}[jac_layout]
    out_types = (jax.ShapeDtypeStruct(rays.shape, jnp.float64),
                 jax.ShapeDtypeStruct((n, max(jdim, 1), max(jdim, 1)), jnp.float64))
    return jax.ffi.ffi_call("tg_trace", out_types, vmap_method="sequential")(
        rays, model=model_bytes, jac_layout=jac_layout)


def make_run_to_end(model):
    """-> ``f(rays (7, N)) -> (7, N)`` with a JVP rule; close over the model like the reference does."""
    mb = _model_bytes(model)

    @jax.custom_jvp
    def run_to_end_soa(rays):
        return _trace(rays, mb, L.TG_JAC_NONE)[0]

    @run_to_end_soa.defjvp
    def _jvp(primals, tangents):
        (rays,), (d_rays,) = primals, tangents
        out, jac = _trace(rays, mb, L.TG_JAC_FULL7)          # (N, 7, 7): d out_i / d in_j per ray
        return out, jnp.einsum("nij,jn->in", jac, d_rays)

    return run_to_end_soa


def run_to_end_abcd(rays, model):
    """``custom_jacobian_matrix(vmap(jacobian(run_to_end))(rays, model))`` in one custom call."""
    return _trace(rays, _model_bytes(model), L.TG_JAC_ABCD5)
