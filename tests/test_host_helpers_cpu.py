"""CPU tests of the host-side (numpy) helpers that sit either side of the accelerated path:
ray samplers / sources (reference utils.py:117-205, source.py:58-188), 5x5 matrix accumulation
(transfer.py:126-183), decompose_Q_inv (gaussian.py:35-89), ParamRef paths."""
import numpy as np
import pytest

from oracle import temgym_oracle as O

from temgymcore_b200.source import ParallelBeam, PointSource
from temgymcore_b200.transfer import accumulate_matrices, accumulate_matrices_cumulative
from temgymcore_b200.utils import concentric_rings, fibonacci_spiral, random_coords


def _reference_rings(n, radius):
    """Literal restatement of utils.py:117-175 incl. the sequential loop of utils.py:46-80."""
    num_rings = max(1, int(np.floor((-1 + np.sqrt(1 + 4 * n / np.pi)) / 2)))
    k = np.round(2 * np.pi * np.arange(1, num_rings + 1)).astype(int)
    ppr = np.round(k * (n / k.sum())).astype(int)
    radii = np.linspace(0, radius, num_rings + 1, endpoint=True)[1:]
    allp = np.repeat(np.stack((radii, 2 * np.pi / ppr), axis=0), ppr.tolist(), axis=-1)
    v = allp[1, :]
    part, cur, cnt = 0, ppr[0], 0
    v[0] = 0.0
    for i in range(1, v.size):
        if cur == cnt:
            cnt, part = 0, part + 1
            cur = ppr[part]
            v[i] = 0.0
        else:
            v[i] += v[i - 1]
            cnt += 1
    return np.stack((allp[0] * np.sin(v), allp[0] * np.cos(v)), axis=-1)


@pytest.mark.parametrize("n", [1, 7, 30, 256, 5000])
def test_concentric_rings_bitwise(n):
    np.testing.assert_array_equal(concentric_rings(n, 0.01), _reference_rings(n, 0.01))


def test_sources_make_rays():
    rays = PointSource(z=0.0, semi_conv=0.01).make_rays(num=256, random=False)  # README.md:272-284
    assert rays.x.shape == rays.dx.shape and rays.size if False else True
    assert rays.z == 0.0 and rays.pathlength == 0.0 and np.all(rays.x == 0) and np.abs(rays.dx).max() <= 0.01
    assert np.hypot(rays.dx, rays.dy).max() <= 0.01 * (1 + 1e-12)
    beam = ParallelBeam(z=1.5, radius=2e-3, offset_xy=(1e-3, -1e-3)).make_rays(500)
    assert np.all(beam.dx == 0) and np.hypot(beam.x - 1e-3, beam.y + 1e-3).max() <= 2e-3 * (1 + 1e-12)
    one = ParallelBeam(z=0.0, radius=0.0).make_rays(1)
    assert np.ndim(one.x) == 0                      # a single ray has scalar fields (source.py:73-78)
    rnd = PointSource(z=0.0, semi_conv=0.02).make_rays(100, random=True)
    assert np.hypot(rnd.dx, rnd.dy).max() < 0.02 and random_coords(0).shape[0] >= 1
    x, y = fibonacci_spiral(1000, 1e-7, alpha=0)
    assert x.shape == (1000,) and np.hypot(x, y).max() <= 1e-7 * (1 + 1e-12) and x[0] == 0.0


def test_accumulate_matrices():
    # reference tests/test_transfer.py:136-153
    A, B, C = np.eye(5), np.eye(5) * 2, np.eye(5) * 3
    for M in (A, B, C):
        M[4, :] = [0, 0, 0, 0, 1]
    np.testing.assert_allclose(accumulate_matrices([A, B, C]), C @ B @ A)
    np.testing.assert_allclose(accumulate_matrices([C]), C)
    rng = np.random.default_rng(0)
    Ms = rng.normal(size=(4, 5, 5))
    cum = accumulate_matrices_cumulative(Ms)
    # the reference's loop order (transfer.py:174-183)
    np.testing.assert_array_equal(cum[0], Ms[3])
    np.testing.assert_array_equal(cum[1], Ms[2] @ Ms[3])
    np.testing.assert_array_equal(cum[3], Ms[0] @ (Ms[1] @ (Ms[2] @ Ms[3])))
    from oracle import temgym_oracle as O
    np.testing.assert_array_equal(cum, O.accumulate_matrices_cumulative(Ms))


def test_decompose_q_inv_parameters():
    from oracle import temgym_oracle as O
    from temgymcore_b200.gaussian import decompose_Q_inv
    wl = 1.3e-6
    for w1, w2, R1, R2, th in [(1e-5, 3e-5, np.inf, np.inf, np.pi / 6), (1.2e-4, 5e-5, 1e-4, 0.1, -np.pi / 4),
                               (1e-4, 3e-4, 0.01, np.inf, -np.pi / 3)]:
        Q = O.gaussian_Q_inv([[w1, w2]], [[R1, R2]], [wl], [th])
        ow1, ow2, oR1, oR2, oth = decompose_Q_inv(Q, wl)
        assert ow1[0] >= ow2[0]
        np.testing.assert_allclose(sorted([ow1[0], ow2[0]]), sorted([w1, w2]), rtol=1e-9)
        # rebuilding Q_inv from the decomposition reproduces its imaginary (envelope) part;
        # the reference reads radii as 1/Re(q) (sign flipped w.r.t. q = -1/R + i...), tested on |field| only
        Q2 = O.gaussian_Q_inv([[ow1[0], ow2[0]]], [[-oR1[0], -oR2[0]]], [wl], [oth[0]])
        np.testing.assert_allclose(Q2, Q, rtol=1e-9, atol=1e-9 * np.abs(Q).max())


def test_param_refs():
    from temgymcore_b200.components import Descanner, DescanError, Lens
    from temgymcore_b200.ray import Ray
    lens = Lens(z=0.5, focal_length=1.0)
    ref = lens.params.focal_length
    assert ref._build() == (lens, "focal_length") and ref._resolve() == 1.0 and ref._resolve_root() is lens
    d = Descanner(z=0.0, scan_pos_x=1.5, scan_pos_y=-2.0, descan_error=DescanError(*np.arange(12.0)))
    assert d.params.descan_error.pxo_pyi._resolve() == 1.0
    assert d._tg_param_seeds(("descan_error", "pxo_pyi")) == [(1, -2.0)]
    assert d._tg_param_seeds(("scan_pos_x",)) == [(1, 0.0 - 1.0), (2, 2.0), (3, 4.0), (4, 6.0)]
    r = Ray(0.0, 0.0, 0.0, 0.0, 0.0, 0.0)
    assert r.params._one._build() == (r, "_one")
    assert {(r, "x"): 1}[(r, "x")] == 1          # rays hash by identity (usable as dict keys)
    assert lens.new_with(focal_length=2.0).focal_length == 2.0


def test_oracle_restates_ring_sampler_and_decomposition():
    """The oracle's concentric_rings / decompose_Q_inv (restatements of utils.py:117-175 and gaussian.py:35-89,
    used to check the device versions) agree with the package's host versions: the ring sampler bit for bit
    (including multi_cumsum_inplace's lagged restarts), the decomposition to rounding."""
    from oracle import temgym_oracle as O
    from temgymcore_b200 import _lib as L
    from temgymcore_b200.gaussian import decompose_Q_inv
    from temgymcore_b200.utils import concentric_rings
    for n in (1, 2, 7, 50, 1234, 20000):
        ref = O.concentric_rings(n, 2.5)
        np.testing.assert_array_equal(concentric_rings(n, 2.5), ref)
        assert L.load().tg_concentric_rings_count(n, 2.5) == ref.shape[0]
    rng = np.random.default_rng(3)
    n = 300
    wl = rng.uniform(1e-12, 5e-12, n)
    Q = O.gaussian_Q_inv(rng.uniform(0.5e-9, 3e-9, (n, 2)), rng.uniform(-1e-3, 1e-3, (n, 2)), wl,
                         rng.uniform(-1.5, 1.5, n))
    a, b = O.decompose_Q_inv(Q, wl[:, None]), decompose_Q_inv(Q, wl[:, None])   # broadcasts like the reference
    for x, y in zip(a, b):
        np.testing.assert_allclose(x, y, rtol=1e-12, atol=1e-15)


def test_container_param_refs_expand_into_leaves():
    """``descanner.params.descan_error`` / ``lens.params.coeffs`` stand for all their leaves, keyed by leaf index
    (the reference's PathBuilder._find_in, tree_utils.py:100-122)."""
    from temgymcore_b200.aberrations import KrivanekCoeffs
    from temgymcore_b200.components import AberratedLensKrivanek, Descanner, DescanError
    from temgymcore_b200.run import _expand_param_leaves
    d = Descanner(z=0.0, scan_pos_x=1.5, scan_pos_y=-2.0, descan_error=DescanError(*np.arange(12.0)))
    leaves = _expand_param_leaves(d, ("descan_error",))
    assert [s for s, _ in leaves] == [(i,) for i in range(12)]
    assert leaves[1][1] == ("descan_error", "pxo_pyi")
    assert all(len(d._tg_param_seeds(leaf)) == 1 for _, leaf in leaves)      # every leaf reaches the kernel
    assert _expand_param_leaves(d, ("scan_pos_x",)) == [((), ("scan_pos_x",))]
    lens = AberratedLensKrivanek(z=0.0, focal_length=1e-3, coeffs=KrivanekCoeffs(C30=1.0))
    cl = _expand_param_leaves(lens, ("coeffs",))
    assert len(cl) == 25 and cl[0][0] == (0,) and all(lens._tg_param_seeds(leaf) for _, leaf in cl)
    assert _expand_param_leaves(lens, ("nope",)) == [((), ("nope",))]       # _tg_param_seeds raises for it


def test_reference_utils_helpers():
    """Host-side helpers of the reference's utils.py / gaussian.py that sit beside the hot path."""
    from temgymcore_b200 import utils as U
    from temgymcore_b200.gaussian import (matrix_linear_mul, matrix_matrix_matrix_mul, matrix_matrix_mul,
                                          matrix_quadratic_mul, matrix_vector_mul)
    # multi_cumsum_inplace: the reference's restart rule (utils.py:69-80): partition k restarts k elements late
    v = np.ones(9)
    U.multi_cumsum_inplace(v, np.array([2, 3, 4]), 0.0)
    np.testing.assert_array_equal(v, [0, 1, 2, 0, 1, 2, 3, 0, 1])
    # inplace_sum: mask + bounds (utils.py:83-114)
    buf = np.zeros((3, 3), np.float32)
    U.inplace_sum(np.array([0, 1, 5, 1, -1]), np.array([0, 1, 1, 1, 0]), np.array([1, 1, 1, 0, 1], bool),
                  np.array([1, 2, 3, 4, 5], np.float32), buf)
    assert buf[0, 0] == 1 and buf[1, 1] == 2 and buf.sum() == 3
    assert U.try_ravel(3.0) == 3.0 and U.try_ravel(np.zeros((2, 2))).shape == (4,)
    assert U.try_reshape(np.arange(4), np.zeros((2, 2))).shape == (2, 2) and U.try_reshape(np.arange(4), 1.0).shape == (4,)
    # FresnelPropagator: energy conserving, identity at z = 0, Gaussian keeps its centre
    n, L, wl = 128, 2e-3, 500e-9
    x = (np.arange(n) - n // 2) * (L / n)
    X, Y = np.meshgrid(x, x)
    u0 = np.exp(-(X ** 2 + Y ** 2) / (2e-4) ** 2).astype(complex)
    np.testing.assert_allclose(U.FresnelPropagator(u0, L, wl, 0.0), u0, atol=1e-12)
    u1 = U.FresnelPropagator(u0, L, wl, 0.05)
    np.testing.assert_allclose(np.sum(np.abs(u1) ** 2), np.sum(np.abs(u0) ** 2), rtol=1e-10)
    assert np.unravel_index(np.argmax(np.abs(u1)), u1.shape) == (n // 2, n // 2)
    z = U.zero_phase(u1.copy(), n // 2, n // 2)
    assert abs(np.angle(z[n // 2, n // 2])) < 1e-12
    m = U.make_aperture(X, Y, aperture_ratio=0.5)
    assert m[n // 2, n // 2] and not m[0, 0]
    two_f = U.fresnel_lens_imaging_solution(u0, Y, X, L / n, wl, 0.1, 0.05, 0.1)   # 2f-2f imaging of a centred,
    np.testing.assert_allclose(np.abs(two_f), np.abs(u0), atol=0.05)               # symmetric beam: |image| = |object|
    # einsum helpers keep the reference's (per-beamlet) index conventions (gaussian.py:180-222)
    rng = np.random.default_rng(0)
    Mx, v, w = rng.normal(size=(2, 2)), rng.normal(size=2), rng.normal(size=(5, 2))
    np.testing.assert_allclose(matrix_vector_mul(Mx, v), Mx @ v)
    np.testing.assert_allclose(matrix_matrix_mul(Mx, Mx), Mx @ Mx)
    np.testing.assert_allclose(matrix_quadratic_mul(v, Mx), v @ Mx @ v)
    np.testing.assert_allclose(matrix_linear_mul(v, Mx, w), w @ (Mx.T @ v))
    B = rng.normal(size=(3, 2, 2))
    np.testing.assert_allclose(matrix_matrix_matrix_mul(B, B, B), np.einsum("nij,njk,npk->nip", B, B, B))


def test_gaussian_ray_q_inv_property():
    from temgymcore_b200.gaussian import GaussianRay
    n = 4
    wl = np.full(n, 2e-12)
    w = np.array([[1e-9, 2e-9]] * n)
    Rc = np.array([[np.inf, 0.5]] * n)
    g = GaussianRay(x=np.zeros(n), y=np.zeros(n), dx=np.zeros(n), dy=np.zeros(n), z=np.zeros(n),
                    pathlength=np.zeros(n), _one=np.ones(n), amplitude=np.ones(n), waist_xy=w,
                    radii_of_curv=Rc, wavelength=wl, theta=np.zeros(n))
    qx, qy = g.q_inv
    np.testing.assert_allclose(qx, 1j * wl / (np.pi * 1e-18))                       # R = inf: purely imaginary
    np.testing.assert_allclose(qy, -2.0 + 1j * wl / (np.pi * 4e-18))
    ox, oy = O.gaussian_q_inv(w, Rc, wl)
    np.testing.assert_array_equal(qx, ox)
    np.testing.assert_array_equal(qy, oy)


def test_fresnel_validators_torch_path_matches_numpy():
    """FresnelPropagator / fresnel_lens_imaging_solution (utils.py:248-275) accept torch tensors (CPU here,
    CUDA on the GPU box: torch.fft = cuFFT) and give the numpy result."""
    import torch
    from temgymcore_b200.utils import FresnelPropagator, fresnel_lens_imaging_solution
    rng = np.random.default_rng(0)
    u = rng.normal(size=(64, 48)) + 1j * rng.normal(size=(64, 48))
    a = FresnelPropagator(u, 1e-3, 5e-7, 0.02)
    b = FresnelPropagator(torch.as_tensor(u), 1e-3, 5e-7, 0.02).numpy()
    np.testing.assert_allclose(b, a, rtol=0, atol=1e-13 * np.abs(a).max())
    u = rng.normal(size=(64, 64)) + 1j * rng.normal(size=(64, 64))
    y, x = np.meshgrid(np.linspace(-1, 1, 64) * 5e-4, np.linspace(-1, 1, 64) * 5e-4, indexing="ij")
    a = fresnel_lens_imaging_solution(u, y, x, 1e-3 / 64, 5e-7, 0.05, 0.03, 0.07)
    b = fresnel_lens_imaging_solution(torch.as_tensor(u), y, x, 1e-3 / 64, 5e-7, 0.05, 0.03, 0.07).numpy()
    np.testing.assert_allclose(b, a, rtol=0, atol=1e-13 * np.abs(a).max())
