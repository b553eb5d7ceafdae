"""The reference's remaining unit-test behaviours (tests/test_component.py, test_rays.py, test_transfer.py,
test_coordinates.py), restated against this package's call surface: components and the propagator are
callable on a Ray exactly like the reference's (``Scanner(...)(ray)``, ``FreeSpaceParaxial()(ray, d)``)
and every value below comes out of the CUDA kernels.  Seeded instead of ``np.random`` at import time."""
import numpy as np
import pytest

from tests import models as M
from temgymcore_b200 import CoordXY, PixelYX
from temgymcore_b200.components import (Descanner, DescanError, Lens, Plane, Scanner)
from temgymcore_b200.propagator import FreeSpaceParaxial
from temgymcore_b200.ray import RAY_FIELDS, Ray

RNG = np.random.default_rng(M.SEED)


@pytest.fixture(scope="module")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def rand_ray(**kw):
    x, y, dx, dy, z, pl = RNG.uniform(-5.0, 5.0, 6)
    d = dict(x=x, y=y, dx=dx, dy=dy, z=z, pathlength=pl, _one=1.0)
    d.update(kw)
    return Ray(**d)


@pytest.mark.gpu
def test_plane_is_identity(gpu):                       # test_component.py:177-183
    ray = rand_ray(z=0.0, pathlength=0.0)
    out = Plane(z=23)(ray)
    for f in RAY_FIELDS:
        assert getattr(out, f) == getattr(ray, f)


@pytest.mark.gpu
@pytest.mark.parametrize("rep", range(3))
def test_scanner_adds_scan_position_and_tilt(gpu, rep):   # test_component.py:186-211
    spx, spy = RNG.uniform(-5, 5, 2)
    stx, sty = RNG.uniform(-0.5, 0.5, 2)
    ray = rand_ray()
    out = Scanner(z=23, scan_pos_x=spx, scan_pos_y=spy, scan_tilt_x=stx, scan_tilt_y=sty)(ray)
    assert (out.x, out.y, out.dx, out.dy) == (ray.x + spx, ray.y + spy, ray.dx + stx, ray.dy + sty)
    assert (out._one, out.z, out.pathlength) == (ray._one, ray.z, ray.pathlength)


@pytest.mark.gpu
@pytest.mark.parametrize("rep", range(3))
def test_descanner_formulas(gpu, rep):                  # test_component.py:214-310
    spx, spy = RNG.uniform(-5, 5, 2)
    stx, sty = RNG.uniform(-0.5, 0.5, 2)
    e = DescanError(*RNG.uniform(0, 1, 12))
    ray = rand_ray()
    out = Descanner(z=23, scan_pos_x=spx, scan_pos_y=spy, scan_tilt_x=stx, scan_tilt_y=sty, descan_error=e)(ray)
    np.testing.assert_allclose(out.x, ray.x + spx * e.pxo_pxi + spy * e.pxo_pyi + e.offpxi - spx, rtol=0, atol=1e-14)
    np.testing.assert_allclose(out.y, ray.y + spx * e.pyo_pxi + spy * e.pyo_pyi + e.offpyi - spy, rtol=0, atol=1e-14)
    np.testing.assert_allclose(out.dx, ray.dx + spx * e.sxo_pxi + spy * e.sxo_pyi + e.offsxi - stx, rtol=0, atol=1e-14)
    np.testing.assert_allclose(out.dy, ray.dy + spx * e.syo_pxi + spy * e.syo_pyi + e.offsyi - sty, rtol=0, atol=1e-14)
    assert (out._one, out.z, out.pathlength) == (ray._one, ray.z, ray.pathlength)
    # a perfect descanner (zero error) undoes the scanner exactly
    sc = Scanner(z=1.0, scan_pos_x=spx, scan_pos_y=spy, scan_tilt_x=stx, scan_tilt_y=sty)
    ds = Descanner(z=1.0, scan_pos_x=spx, scan_pos_y=spy, scan_tilt_x=stx, scan_tilt_y=sty, descan_error=DescanError())
    back = ds(sc(ray))
    np.testing.assert_allclose([back.x, back.y, back.dx, back.dy], [ray.x, ray.y, ray.dx, ray.dy], rtol=0, atol=1e-14)


@pytest.mark.gpu
def test_singular_jacobian_has_no_finite_inverse(gpu):   # test_component.py:381-395
    from temgymcore_b200.run import run_to_end_abcd
    # the reference builds a component whose Jacobian is singular and checks that inverting the 5x5
    # yields nan / inf rather than an exception; a lens of zero focal length does the same here
    ray = Ray(x=0.0, y=0.0, dx=1.0, dy=1.0, z=0.0, pathlength=0.0)
    _, J = run_to_end_abcd(ray, [Lens(z=0.0, focal_length=0.0)])     # -x / 0 -> inf / nan entries
    assert not np.isfinite(J).all()
    with np.errstate(all="ignore"):
        try:
            inv = np.linalg.inv(J)
            assert np.isnan(inv).any() or np.isinf(inv).any()
        except np.linalg.LinAlgError:
            pass


@pytest.mark.gpu
def test_propagator_call(gpu):                            # test_rays.py:15-39
    ray = Ray(x=1.0, y=-1.0, dx=0.5, dy=0.5, z=2.0, pathlength=1.0)
    same = FreeSpaceParaxial()(ray, 0.0)
    for f in RAY_FIELDS:
        assert getattr(same, f) == getattr(ray, f)
    for _ in range(5):
        d = RNG.uniform(0.1, 10.0)
        r = rand_ray()
        new = FreeSpaceParaxial()(r, d)
        assert (new.x, new.y, new.dx, new.dy, new.z, new.pathlength) == (
            r.x + r.dx * d, r.y + r.dy * d, r.dx, r.dy, r.z + d, r.pathlength + d)


@pytest.mark.gpu
def test_propagation_z_does_not_depend_on_one(gpu):       # test_rays.py:42-57
    from temgymcore_b200.run import ray_jacobian
    ray = Ray(x=0.5, y=-0.5, dx=0.1, dy=-0.2, z=0.0, pathlength=0.0, _one=1.0)
    jac = ray_jacobian(ray, [Plane(z=0.1)])
    assert jac.z._one == 0.0 and jac.pathlength._one == 0.0
    assert jac.x.dx == 0.1 and jac.x.x == 1.0


def test_to_vector_and_namedtuples():                     # test_rays.py:106-114, test_coordinates.py
    ray = Ray(x=1.0, y=2.0, dx=0.1, dy=0.2, z=3.0, pathlength=4.0)
    v = ray.to_vector()
    for f in RAY_FIELDS:
        a = getattr(v, f)
        assert isinstance(a, np.ndarray) and a.shape == (1,)
    p = PixelYX(y=3, x=5).to_pixels()
    assert p.y.tolist() == [3] and p.x.tolist() == [5] and np.issubdtype(p.y.dtype, np.integer)
    pf = PixelYX(y=3.5, x=5.25).to_pixels()
    assert np.issubdtype(pf.x.dtype, np.floating)
    c = CoordXY(x=1.5, y=-2.0).to_coords()
    assert c.x.tolist() == [1.5] and c.y.tolist() == [-2.0]


@pytest.mark.gpu
def test_transfer_pt_src(gpu):                            # test_transfer.py:7-117
    from temgymcore_b200.transfer import transfer_rays_pt_src

    def drift(d):
        T = np.eye(5)
        T[0, 2] = T[1, 3] = d
        return T
    for d in (5.0, -3.0, 0.0):
        sx, sy = np.array([np.cos(0.7)]), np.array([np.sin(0.7)])
        out = transfer_rays_pt_src((1.0, -2.0), (sx, sy), drift(d))
        np.testing.assert_allclose(out[:, 0], [1.0 + d * sx[0], -2.0 + d * sy[0], sx[0], sy[0]], rtol=0, atol=1e-15)
    T = RNG.integers(-5, 5, (5, 5)).astype(float)
    T[4] = [0, 0, 0, 0, 1]
    out = transfer_rays_pt_src((0.7, -1.2), (np.array([0.3]), np.array([-0.4])), T)
    np.testing.assert_allclose(out[:, 0], (T @ np.array([0.7, -1.2, 0.3, -0.4, 1.0]))[:4], rtol=0, atol=1e-14)
    dxs, dys = RNG.standard_normal(4), RNG.standard_normal(4)
    out = transfer_rays_pt_src((0.5, -0.5), (dxs, dys), np.eye(5))
    assert out.shape == (4, 4)
    np.testing.assert_array_equal(out, [np.full(4, 0.5), np.full(4, -0.5), dxs, dys])
    assert transfer_rays_pt_src((1.0, 2.0), (np.array([]), np.array([])), np.eye(5)).shape == (4, 0)
