"""Wave-optics cross-checks in the style of the reference's own tests/test_gaussians.py: the beamlet
field (oracle on the CPU, CUDA path on the GPU) against an independent FFT Fresnel propagation of the
input-plane field.  These are the only reference tests that exercise MISALIGNED beamlets (offset and
tilted central rays); their tolerances are loose by construction (sampling of the FFT propagator).

  * free space vs Fresnel, amplitude + phase inside an aperture   (test_gaussians.py:276-371, rtol/atol 0.1)
  * tilted beam lands in the right quadrant                        (test_gaussians.py:684-729)
  * random offset + tilt: peak position vs Fresnel within 3 pixels  (test_gaussians.py:788-817)
"""
import numpy as np
import pytest

from oracle import temgym_oracle as O
from tests import models as M
from temgymcore_b200.components import Detector


def fresnel_transfer(u0, width, wavelength, z):
    """Paraxial free-space propagation of a sampled field by the transfer-function method (the
    reference's validator, utils.py:248-265): multiply the spectrum by exp(-i pi lambda z |f|^2)."""
    ny, nx = u0.shape
    step = width / ny
    fx = np.fft.fftfreq(nx, d=step)
    fy = np.fft.fftfreq(ny, d=step)
    f2 = fx[None, :] ** 2 + fy[:, None] ** 2
    return np.fft.ifft2(np.fft.fft2(u0) * np.exp(-1j * np.pi * wavelength * z * f2))


def single_beamlet(x0, y0, dx, dy, wavelength=500e-9, w0=1e-4):
    return M.gaussian_rays([x0], [y0], dx=[dx], dy=[dy], wavelength=wavelength, w0=w0)


def det_axes(det):
    ny, nx = det.shape
    xs = (np.arange(nx) - (nx - 1) / 2) * det.pixel_size[1]
    ys = -((np.arange(ny) - (ny - 1) / 2) * det.pixel_size[0])
    return np.meshgrid(xs, ys)       # X, Y with Y decreasing down the rows (grid.py / SURVEY A.5)


class OracleBackend:
    name = "oracle"

    @staticmethod
    def image(g, model):
        return O.make_gaussian_image(g, model)

    @staticmethod
    def input_image(g, det):
        return O.evaluate_gaussian_input_image(g, det)


class CudaBackend:
    name = "cuda"

    @staticmethod
    def image(g, model):
        from temgymcore_b200.gaussian import make_gaussian_image
        return np.asarray(make_gaussian_image(g, model, cull_bits=0))

    @staticmethod
    def input_image(g, det):
        from temgymcore_b200.gaussian import evaluate_gaussian_input_image
        return np.asarray(evaluate_gaussian_input_image(g, det, cull_bits=0))


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request):
    if request.param == "cuda":
        import torch
        if not torch.cuda.is_available():
            pytest.skip("no CUDA device")
        return CudaBackend
    return OracleBackend


def test_free_space_vs_fresnel(backend):
    wavelength, w0, z = 1e-3, 0.1, 1e-3
    n = 500
    det = Detector(z=z, pixel_size=(2e-3, 2e-3), shape=(n, n))
    g = M.gaussian_rays([0.0], [0.0], wavelength=wavelength, w0=w0)
    X, Y = det_axes(det)
    analytic = backend.image(g, [det])
    u0 = np.exp(-(X ** 2 + Y ** 2) / w0 ** 2).astype(complex)     # waist at z = 0: flat phase
    fres = fresnel_transfer(u0, n * det.pixel_size[0], wavelength, z)
    analytic = analytic / np.abs(analytic).max()
    fres = fres / np.abs(fres).max()
    mask = X ** 2 + Y ** 2 < (0.4 * np.abs(X).max()) ** 2
    np.testing.assert_allclose(np.abs(analytic)[mask], np.abs(fres)[mask], rtol=1e-1, atol=1e-2)
    # phase relative to the centre pixel, inside the aperture where the amplitude is significant
    c = n // 2
    ph_a = np.angle(analytic * np.conj(analytic[c, c]))
    ph_f = np.angle(fres * np.conj(fres[c, c]))
    strong = mask & (np.abs(fres) > 0.05)
    assert np.abs(np.angle(np.exp(1j * (ph_a - ph_f))))[strong].max() < 1e-1


def test_tilted_beam_lands_in_the_right_quadrant(backend):
    det = Detector(z=1.0, pixel_size=(1e-4, 1e-4), shape=(128, 128))
    X, Y = det_axes(det)
    ref = backend.input_image(single_beamlet(0.0, 0.0, 0.0, 0.0), det)
    cy, cx = np.unravel_index(np.argmax(np.abs(ref)), ref.shape)
    t = 1e-3
    for dx, dy in ((t, t), (t, -t), (-t, t), (-t, -t)):
        img = backend.image(single_beamlet(0.0, 0.0, dx, dy), [det])
        my, mx = np.unravel_index(np.argmax(np.abs(img)), img.shape)
        assert np.sign(X[my, mx] - X[cy, cx]) == np.sign(dx)
        assert np.sign(Y[my, mx] - Y[cy, cx]) == np.sign(dy)


def test_offset_tilted_beam_vs_fresnel(backend):
    rng = np.random.default_rng(M.SEED)
    n = 512
    det = Detector(z=1.0, pixel_size=(2e-5, 2e-5), shape=(n, n))
    for _ in range(3):
        r1m = rng.uniform(-1e-3, 1e-3, 2)
        th = rng.uniform(-1e-4, 1e-4, 2)
        g = single_beamlet(r1m[0], r1m[1], th[0], th[1])
        u0 = backend.input_image(g, det)
        out = backend.image(g, [det])
        fres = fresnel_transfer(u0, n * det.pixel_size[0], 500e-9, 1.0)
        pf = np.unravel_index(np.argmax(np.abs(fres)), fres.shape)
        po = np.unravel_index(np.argmax(np.abs(out)), out.shape)
        assert abs(pf[0] - po[0]) <= 3 and abs(pf[1] - po[1]) <= 3
        # stronger than the reference's check: the whole normalised amplitude profile agrees
        a, b = np.abs(out) / np.abs(out).max(), np.abs(fres) / np.abs(fres).max()
        assert np.abs(a - b).max() < 0.05


def test_astigmatic_rotated_beam_axis(backend):
    """An elliptical beamlet (w_x = 2 w_y) rotated by theta: the major axis of its input-plane intensity
    (second moments) points along theta (test_gaussians.py:820-883, tolerance 0.1 rad there; the
    moments are exact to ~1e-3 rad here).  After free-space propagation the far field of the NARROW axis
    spreads faster, so the major axis has turned by 90 degrees -- checked through make_gaussian_image,
    which the reference's test does not do."""
    det = Detector(z=0.0, pixel_size=(2e-5, 2e-5), shape=(256, 256))
    X, Y = det_axes(det)

    def major_axis(field):
        I = np.abs(field) ** 2
        w = I.sum()
        xc, yc = X - (I * X).sum() / w, Y - (I * Y).sum() / w
        C = np.array([[(I * xc * xc).sum(), (I * xc * yc).sum()], [(I * xc * yc).sum(), (I * yc * yc).sum()]]) / w
        vals, vecs = np.linalg.eigh(C)
        v = vecs[:, np.argmax(vals)]
        return (np.arctan2(v[1], v[0]) + np.pi / 2) % np.pi - np.pi / 2

    def axis_diff(a, b):
        d = np.arctan2(np.sin(a - b), np.cos(a - b))
        return min(abs(d), abs(d + np.pi), abs(d - np.pi))

    for theta in (np.pi / 3, np.pi / 6, 0.0, -np.pi / 6, -np.pi / 3, -np.pi / 2 + 0.1):
        g = M.gaussian_rays([0.0], [0.0], wavelength=500e-9, waist_xy=[[2e-4, 1e-4]], theta=[theta])
        img = backend.input_image(g, det)
        assert axis_diff(major_axis(img), theta) < 5e-3
        far = Detector(z=2.0, pixel_size=(8e-5, 8e-5), shape=(256, 256))     # z >> z_R of both axes
        Xf, Yf = det_axes(far)
        out = backend.image(g, [far])
        I = np.abs(out) ** 2
        w = I.sum()
        C = np.array([[(I * Xf * Xf).sum(), (I * Xf * Yf).sum()], [(I * Xf * Yf).sum(), (I * Yf * Yf).sum()]]) / w
        vals, vecs = np.linalg.eigh(C)
        v = vecs[:, np.argmax(vals)]
        assert axis_diff(np.arctan2(v[1], v[0]), theta + np.pi / 2) < 2e-2


def test_offset_tilted_beam_through_lens_vs_fresnel(backend):
    """A misaligned beamlet (offset AND tilted central ray) through a thin lens onto a detector that is
    NOT in the image plane: every block of the ABCD matrix is non-trivial (A = 1 - z2/f, B = z1 + z2 -
    z1 z2/f, C = -1/f, D = 1 - z1/f).  Compared with Fresnel propagation -> lens phase -> Fresnel
    propagation of the sampled input field (the reference's fresnel_lens_imaging_solution validator,
    utils.py:268-275; its own lens test, test_gaussians.py:374-518, is on-axis with rtol 0.5)."""
    from temgymcore_b200.components import Lens
    from temgymcore_b200.utils import fresnel_lens_imaging_solution
    wl, w0 = 500e-9, 2e-4
    z1, f, z2 = 0.3, 0.2, 0.25
    n, px = 512, 1e-5
    det_in = Detector(z=0.0, pixel_size=(px, px), shape=(n, n))
    det_out = Detector(z=z1 + z2, pixel_size=(px, px), shape=(n, n))
    X, Y = det_axes(det_in)
    rng = np.random.default_rng(M.SEED + 3)
    for _ in range(2):
        r1m = rng.uniform(-3e-4, 3e-4, 2)
        th = rng.uniform(-2e-4, 2e-4, 2)
        g = M.gaussian_rays([r1m[0]], [r1m[1]], dx=[th[0]], dy=[th[1]], wavelength=wl, w0=w0)
        u0 = backend.input_image(g, det_in)
        ref = fresnel_lens_imaging_solution(u0, Y, X, px, wl, z1, f, z2)
        out = backend.image(g, [Lens(z=z1, focal_length=f), det_out])
        pa = np.unravel_index(np.argmax(np.abs(out)), out.shape)
        pf = np.unravel_index(np.argmax(np.abs(ref)), ref.shape)
        assert abs(pa[0] - pf[0]) <= 2 and abs(pa[1] - pf[1]) <= 2
        # the ray-optics prediction of the spot centre agrees too: x_out = A x + B dx
        A_, B_ = 1 - z2 / f, z1 + z2 - z1 * z2 / f
        assert abs(X[pa] - (A_ * r1m[0] + B_ * th[0])) <= 2 * px and abs(Y[pa] - (A_ * r1m[1] + B_ * th[1])) <= 2 * px
        a, b = np.abs(out) / np.abs(out).max(), np.abs(ref) / np.abs(ref).max()
        assert np.abs(a - b).max() < 0.03
        # phase relative to the peak pixel where the beam is bright
        strong = b > 0.3
        dphi = np.angle(out * np.conj(out[pf]) * np.conj(ref * np.conj(ref[pf])))
        assert np.abs(dphi[strong]).max() < 0.1


@pytest.mark.gpu
def test_device_side_fresnel_cross_check():
    """the wave-optics cross-check without leaving the GPU -- the beamlet image comes back
    as a CUDA tensor, the validator (FresnelPropagator on torch.fft / cuFFT) runs on the same device."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from dataclasses import fields, replace
    from temgymcore_b200.gaussian import make_gaussian_image_device
    from temgymcore_b200.utils import FresnelPropagator
    wavelength, w0, z = 1e-3, 0.1, 1e-3
    n = 500
    det = Detector(z=z, pixel_size=(2e-3, 2e-3), shape=(n, n))
    g = M.gaussian_rays([0.0], [0.0], wavelength=wavelength, w0=w0)
    gd = replace(g, **{f.name: torch.as_tensor(np.asarray(getattr(g, f.name)), device="cuda") for f in fields(g)})
    analytic = make_gaussian_image_device(gd, [det], cull_bits=0)
    assert analytic.is_cuda
    X, Y = det_axes(det)
    u0 = torch.as_tensor(np.exp(-(X ** 2 + Y ** 2) / w0 ** 2).astype(complex), device="cuda")
    fres = FresnelPropagator(u0, n * det.pixel_size[0], wavelength, z)
    assert fres.is_cuda
    np.testing.assert_allclose(fres.cpu().numpy(), fresnel_transfer(u0.cpu().numpy(), n * det.pixel_size[0],
                                                                  wavelength, z), atol=1e-12)
    a = analytic / analytic.abs().max()
    f_ = fres / fres.abs().max()
    mask = torch.as_tensor(X ** 2 + Y ** 2 < (0.4 * np.abs(X).max()) ** 2, device="cuda")
    assert float((a.abs() - f_.abs())[mask].abs().max()) < 2e-2
