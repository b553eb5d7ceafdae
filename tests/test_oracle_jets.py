"""Pins the oracle's higher-order derivative engine (truncated-Taylor ``Jet`` algebra restating the
reference's nested ``jax.jacfwd``, run.py:119-147) against sympy's symbolic differentiation -- the
reference has no test or golden for ``calculate_derivatives``."""
import itertools

import numpy as np
import pytest
import sympy as sp

from oracle import temgym_oracle as O
from tests import models as M

SYMS = sp.symbols("x y dx dy z pl one", real=True)


def _check(exprs, derivs, point, rng, n_samples, rtol=1e-11, outputs=range(7)):
    """exprs: 7 sympy expressions (outputs in RAY_FIELDS order); derivs: oracle tensors for ONE ray.
    A random sample of index tuples per order; the variables a tuple does not differentiate are
    substituted numerically (30 digits) BEFORE differentiating, which keeps sympy fast."""
    for order in (1, 2, 3):
        D = derivs[order - 1][0]
        scale = np.abs(D).max() or 1.0
        combos = list(itertools.combinations_with_replacement(range(7), order))
        picks = [combos[i] for i in rng.choice(len(combos), size=min(n_samples, len(combos)), replace=False)]
        for idx in picks:
            keep = set(idx)
            fixed = {SYMS[i]: sp.Float(point[i], 30) for i in range(7) if i not in keep}
            at = {SYMS[i]: sp.Float(point[i], 30) for i in keep}
            for f in outputs:
                fscale = np.abs(D[f]).max() or scale
                small = exprs[f].subs(fixed)
                want = float(sp.diff(small, *[SYMS[i] for i in idx]).subs(at).evalf(25))
                for perm in set(itertools.permutations(idx)):
                    got = D[(f,) + perm]
                    assert abs(got - want) <= rtol * max(fscale, abs(want)), (order, f, perm, got, want)


def test_jets_rational_model_vs_sympy():
    """Free space, Lens, Deflector, Scanner, Rotator, ThickLens: the oracle's own component code run on
    sympy symbols gives closed forms; their symbolic derivatives must equal the Jet coefficients."""
    model = [c for c in M.kitchen_sink_model() if type(c).__name__ not in ("Biprism",)]
    point = [0.013, -0.021, 0.004, -0.007, -0.15, 0.3, 1.0]
    sym_out = O.run_to_end(O.Ray(*SYMS), model)
    exprs = [sp.sympify(getattr(sym_out, f)) for f in O.RAY_FIELDS]
    ray = O.Ray(*[np.array([v]) for v in point])
    derivs = O.calculate_derivatives(ray, model, 3)
    _check(exprs, derivs, point, np.random.default_rng(0), 12)
    # first order == the forward-mode dual Jacobian the other oracle tests pin
    np.testing.assert_allclose(derivs[0], O.jacobian_run_to_end(ray, model)[1], rtol=1e-13, atol=1e-15)


def _krivanek_symbolic(x, y, dx, dy, pl, f, C):
    """AberratedLensKrivanek written out by hand from components.py:192-215 / aberrations.py:34-108
    for the coefficients C10, C12/phi12, C21/phi21, C23/phi23, C30 (sympy atan2 / sqrt / cos / sin)."""
    ix, iy = -x / f + dx, -y / f + dy
    a = sp.sqrt(ix ** 2 + iy ** 2)
    ph = sp.atan2(iy, ix)
    B2 = C["C10"] + C["C12"] * sp.cos(2 * (ph - C["phi12"]))
    B3 = C["C21"] * sp.cos(ph - C["phi21"]) + C["C23"] * sp.cos(3 * (ph - C["phi23"]))
    B4 = C["C30"]
    W = a ** 2 / 2 * B2 + a ** 3 / 3 * B3 + a ** 4 / 4 * B4
    dWa = a * B2 + a ** 2 * B3 + a ** 3 * B4
    dWp = (a ** 2 / 2) * (-2 * C["C12"] * sp.sin(2 * (ph - C["phi12"]))) + (a ** 3 / 3) * (
        -C["C21"] * sp.sin(ph - C["phi21"]) - 3 * C["C23"] * sp.sin(3 * (ph - C["phi23"])))
    dWx = dWa * (ix / a) + dWp * (-iy / a ** 2)
    dWy = dWa * (iy / a) + dWp * (ix / a ** 2)
    return ix - dWx / f, iy - dWy / f, pl - (x ** 2 + y ** 2) / (2 * f) + W / f


def test_jets_krivanek_vs_sympy():
    from temgymcore_b200.aberrations import KrivanekCoeffs
    from temgymcore_b200.components import AberratedLensKrivanek, Detector
    C = dict(C10=0.3, C12=0.2, phi12=0.4, C21=1.5, phi21=-0.3, C23=0.7, phi23=1.1, C30=2.0)
    f, zl, zd = 0.8, 0.25, 0.6
    model = [AberratedLensKrivanek(z=zl, focal_length=f, coeffs=KrivanekCoeffs(**C)),
             Detector(z=zd, pixel_size=(0.01, 0.01), shape=(8, 8))]
    x, y, dx, dy, z, pl, one = SYMS
    d1 = zl - z
    x1, y1, pl1 = x + dx * d1, y + dy * d1, pl + d1
    ndx, ndy, pl2 = _krivanek_symbolic(x1, y1, dx, dy, pl1, sp.Float(f), {k: sp.Float(v) for k, v in C.items()})
    d2 = zd - zl
    exprs = [x1 + ndx * d2, y1 + ndy * d2, ndx, ndy, sp.Float(zd) + 0 * z, pl2 + d2, one * 1.0]
    point = [0.11, -0.07, 0.23, 0.31, 0.05, 0.0, 1.0]
    ray = O.Ray(*[np.array([v]) for v in point])
    derivs = O.calculate_derivatives(ray, model, 3)
    vals = O.run_to_end(ray, model)
    subs = dict(zip(SYMS, point))
    for i, fld in enumerate(O.RAY_FIELDS):
        assert abs(float(exprs[i].evalf(30, subs=subs)) - float(np.asarray(getattr(vals, fld)).reshape(-1)[0])) < 1e-13
    _check(exprs, derivs, point, np.random.default_rng(1), 6, rtol=1e-10, outputs=(0, 2, 3, 5))
    # the first-order jets are the Krivanek Jacobian of the forward-mode duals (what the ABCD tests use):
    # both restatements agree, and sympy pins them
    J = O.jacobian_run_to_end(ray, model)[1]
    np.testing.assert_allclose(derivs[0], J, rtol=1e-11, atol=1e-13 * np.abs(J).max())


def test_jets_pathlength_input_is_additive():
    """The CUDA kernel treats the pathlength INPUT analytically (d pl_out / d pl_in = 1, everything else
    involving pl_in is 0): true for every component, checked on the oracle's full 7-variable jets."""
    for model, rays in ((M.kitchen_sink_model(), M.random_rays(7)),
                        (M.six_component_column(), M.random_rays(7, scale=0.2e-9, slope=1e-6))):
        d1, d2, d3 = O.calculate_derivatives(rays, model, 3)
        assert np.all(d1[:, 5, 5] == 1.0)
        mask = np.ones(7, bool)
        mask[5] = False
        assert np.all(d1[:, mask, 5] == 0.0)
        assert np.all(d2[:, :, 5, :] == 0.0) and np.all(d2[:, :, :, 5] == 0.0)
        assert np.all(d3[:, :, 5] == 0.0) and np.all(d3[:, :, :, 5] == 0.0) and np.all(d3[..., 5] == 0.0)
