"""The polynomial form of the Krivanek aberration function that the CUDA kernels evaluate (csrc/trace.cu,
csrc/jets.cu), restated in numpy and checked on the CPU against the oracle's polar form (the reference's
formulas, aberrations.py:42-108) -- value, gradient, and the Hessian against central differences.

    w = ax + i ay,  r2 = |w|^2,  b = (n + 1 - m) / 2,  z_nm = C_nm / (n + 1) exp(-i m phi_nm)
    W = Re[ F0(w) + r2 F1(w) + r2^2 F2(w) + r2^3 F3 ],   F_b = sum of z_nm w^m over the terms with that b
"""
from types import SimpleNamespace

import numpy as np
import pytest

from oracle import temgym_oracle as O

# (coefficient, phase, n, m) in KrivanekCoeffs order
TERMS = [("C10", None, 1, 0), ("C12", "phi12", 1, 2), ("C21", "phi21", 2, 1), ("C23", "phi23", 2, 3),
         ("C30", None, 3, 0), ("C32", "phi32", 3, 2), ("C34", "phi34", 3, 4), ("C41", "phi41", 4, 1),
         ("C43", "phi43", 4, 3), ("C45", "phi45", 4, 5), ("C50", None, 5, 0), ("C52", "phi52", 5, 2),
         ("C54", "phi54", 5, 4), ("C56", "phi56", 5, 6)]


def poly_form(p, u, v):
    w = u + 1j * v
    r2 = u * u + v * v
    F = [0, 0, 0, 0]
    F1 = [0, 0, 0, 0]
    F2 = [0, 0, 0, 0]
    for c, ph, n, m in TERMS:
        z = getattr(p, c) / (n + 1) * np.exp(-1j * m * (getattr(p, ph) if ph else 0.0))
        b = (n + 1 - m) // 2
        F[b] = F[b] + z * w ** m
        F1[b] = F1[b] + (m * z * w ** (m - 1) if m >= 1 else 0)
        F2[b] = F2[b] + (m * (m - 1) * z * w ** (m - 2) if m >= 2 else 0)
    W = G = Hxx = Hxy = Hyy = 0
    for b in range(4):
        f, g, f2 = np.real(F[b]), np.conj(F1[b]), F2[b]           # grad Re F = conj(F'), Hess from F''
        rho = r2 ** b
        rho1 = b * r2 ** (b - 1) if b >= 1 else 0
        rho2 = b * (b - 1) * r2 ** (b - 2) if b >= 2 else 0
        W = W + rho * f
        G = G + rho * g + 2 * rho1 * f * w
        Hxx = Hxx + rho * np.real(f2) + 2 * rho1 * f + 4 * rho2 * f * u * u + 4 * rho1 * u * np.real(g)
        Hxy = Hxy - rho * np.imag(f2) + 4 * rho2 * f * u * v + 2 * rho1 * (u * np.imag(g) + v * np.real(g))
        Hyy = Hyy - rho * np.real(f2) + 2 * rho1 * f + 4 * rho2 * f * v * v + 4 * rho1 * v * np.imag(g)
    return W, np.real(G), np.imag(G), Hxx, Hxy, Hyy


def _random_coeffs(rng):
    vals = {f: (rng.uniform(-np.pi, np.pi) if f.startswith("phi") else rng.uniform(-1, 1) * 10 ** rng.uniform(0, 3))
            for f in O._KRIV_FIELDS}
    return SimpleNamespace(**vals)


def test_polynomial_form_equals_polar_form():
    rng = np.random.default_rng(1)
    for _ in range(3):
        p = _random_coeffs(rng)
        u, v = rng.uniform(-1, 1, (2, 2000)) * 1e-2
        W, Gx, Gy, Hxx, Hxy, Hyy = poly_form(p, u, v)
        Wr = O.W_krivanek(np.hypot(u, v), np.arctan2(v, u), p)
        gxr, gyr = O.grad_W_krivanek(u, v, p)
        np.testing.assert_allclose(W, Wr, rtol=1e-11, atol=1e-14 * np.abs(Wr).max())
        np.testing.assert_allclose(Gx, gxr, rtol=1e-10, atol=1e-13 * np.abs(gxr).max())
        np.testing.assert_allclose(Gy, gyr, rtol=1e-10, atol=1e-13 * np.abs(gyr).max())
        h = 1e-7
        gxp, gyp = O.grad_W_krivanek(u + h, v, p)
        gxm, gym = O.grad_W_krivanek(u - h, v, p)
        scale = max(np.abs(Hxx).max(), np.abs(Hxy).max(), np.abs(Hyy).max())
        np.testing.assert_allclose(Hxx, (gxp - gxm) / (2 * h), rtol=1e-6, atol=1e-7 * scale)
        np.testing.assert_allclose(Hxy, (gyp - gym) / (2 * h), rtol=1e-6, atol=1e-7 * scale)
        gxp, gyp = O.grad_W_krivanek(u, v + h, p)
        gxm, gym = O.grad_W_krivanek(u, v - h, p)
        np.testing.assert_allclose(Hxy, (gxp - gxm) / (2 * h), rtol=1e-6, atol=1e-7 * scale)
        np.testing.assert_allclose(Hyy, (gyp - gym) / (2 * h), rtol=1e-6, atol=1e-7 * scale)


def test_one_term_derivatives_of_the_polynomial_form():
    """(W, grad W) is linear in every C_nm, so d/dC_nm is the polynomial of a one-term coefficient set; a phase
    phi_nm only enters through z_nm, whose derivative -i m z_nm is the one-term set with C' = m C and the phase
    advanced by a quarter period -- what krivanek_poly_dfield does for the coefficient tangents of run_with_grads."""
    rng = np.random.default_rng(2)
    p = _random_coeffs(rng)
    u, v = rng.uniform(-1, 1, (2, 500)) * 1e-2
    base = poly_form(p, u, v)[:3]
    scale = [np.abs(b).max() for b in base]
    zero = {f: 0.0 for f in O._KRIV_FIELDS}
    for c, ph, n, m in TERMS:
        delta = abs(getattr(p, c)) + 1.0
        q = SimpleNamespace(**vars(p))
        setattr(q, c, getattr(p, c) + delta)
        one = SimpleNamespace(**zero)
        setattr(one, c, 1.0)
        if ph:
            setattr(one, ph, getattr(p, ph))
        for got, b0, an, sc in zip(poly_form(q, u, v)[:3], base, poly_form(one, u, v)[:3], scale):
            np.testing.assert_allclose(got - b0, delta * an, rtol=0, atol=1e-11 * (sc + delta * np.abs(an).max()))
        if ph:
            h = 1e-6
            qp, qm = SimpleNamespace(**vars(p)), SimpleNamespace(**vars(p))
            setattr(qp, ph, getattr(p, ph) + h)
            setattr(qm, ph, getattr(p, ph) - h)
            dphi = SimpleNamespace(**zero)
            setattr(dphi, c, m * getattr(p, c))
            setattr(dphi, ph, getattr(p, ph) + np.pi / (2 * m))
            for fp, fm, an, sc in zip(poly_form(qp, u, v)[:3], poly_form(qm, u, v)[:3], poly_form(dphi, u, v)[:3], scale):
                np.testing.assert_allclose((fp - fm) / (2 * h), an, rtol=0, atol=1e-8 * sc / h * 1e-6 + 1e-7 * np.abs(an).max())


def wirtinger_table(p, u0, v0, order=4):
    """csrc/jets.cu kriv_table restated: W_ij = d^(i+j) W / du^i dv^j for i + j <= order from the Wirtinger
    derivatives G_pq = sum_terms z ff(a, p) ff(b, q) w^(a-p) conj(w)^(b-q) of G = sum z w^a conj(w)^b, W = Re G."""
    from math import comb

    def ff(a, k):
        r = 1
        for t in range(k):
            r *= a - t
        return r

    w = u0 + 1j * v0
    G = np.zeros((order + 1, order + 1), dtype=complex)
    for c, ph, n, m in TERMS:
        z = getattr(p, c) / (n + 1) * np.exp(-1j * m * (getattr(p, ph) if ph else 0.0))
        b = (n + 1 - m) // 2
        a = m + b
        for pp in range(order + 1):
            for q in range(order + 1 - pp):
                if a >= pp and b >= q:
                    G[pp, q] += z * ff(a, pp) * ff(b, q) * w ** (a - pp) * np.conj(w) ** (b - q)
    T = np.zeros((order + 1, order + 1))
    for i in range(order + 1):
        for j in range(order + 1 - i):
            acc = sum(comb(i, k) * comb(j, l) * (-1) ** (j - l) * G[k + l, (i - k) + (j - l)]
                      for k in range(i + 1) for l in range(j + 1))
            T[i, j] = np.real(1j ** j * acc)
    return T


def test_wirtinger_table_of_partials_vs_sympy():
    """The per-ray table the jet kernel composes with (orders 0..4) against symbolic differentiation of the polar
    form's polynomial, and its low orders against poly_form (value, gradient, Hessian)."""
    sp = pytest.importorskip("sympy")
    rng = np.random.default_rng(5)
    p = _random_coeffs(rng)
    u0, v0 = 0.013, -0.007
    T = wirtinger_table(p, u0, v0)
    W, Gx, Gy, Hxx, Hxy, Hyy = poly_form(p, np.array([u0]), np.array([v0]))
    for got, ref in ((T[0, 0], W), (T[1, 0], Gx), (T[0, 1], Gy), (T[2, 0], Hxx), (T[1, 1], Hxy), (T[0, 2], Hyy)):
        np.testing.assert_allclose(got, ref[0], rtol=1e-12)
    u, v = sp.symbols("u v", real=True)
    Ws = 0
    for c, ph, n, m in TERMS:
        z = sp.Float(getattr(p, c), 30) / (n + 1) * sp.exp(-sp.I * m * sp.Float(getattr(p, ph) if ph else 0.0, 30))
        Ws += sp.re(sp.expand(z * (u + sp.I * v) ** m)) * (u * u + v * v) ** ((n + 1 - m) // 2)
    for i in range(5):
        for j in range(5 - i):
            d = Ws
            if i:
                d = sp.diff(d, u, i)
            if j:
                d = sp.diff(d, v, j)
            ref = float(d.subs({u: sp.Float(u0, 30), v: sp.Float(v0, 30)}))
            assert abs(T[i, j] - ref) <= 1e-12 * abs(ref) + 1e-300, (i, j, T[i, j], ref)
