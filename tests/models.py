"""Model / input builders shared by the CPU (oracle) and GPU (parity) tests.

Models are built from the product's component dataclasses (pure parameter holders);
the oracle duck-types them by class name, so the same objects drive both sides.
"""
import numpy as np

from temgymcore_b200.aberrations import KrivanekCoeffs
from temgymcore_b200.components import (AberratedLensKrivanek, Biprism, Deflector, Descanner,
                                        DescanError, Detector, Lens, Plane, Rotator, ScanGrid,
                                        Scanner, ThickLens)
from temgymcore_b200.gaussian import GaussianRay
from temgymcore_b200.ray import Ray
from temgymcore_b200.source import ParallelBeam, PointSource
from temgymcore_b200.utils import fibonacci_spiral

SEED = 20261017
_CLS = dict(Lens=Lens, Detector=Detector, ParallelBeam=ParallelBeam, Biprism=Biprism,
            Deflector=Deflector, Plane=Plane)


def build(spec):
    out = []
    for name, kw in spec:
        kw = dict(kw)
        for k in ("pixel_size", "shape"):
            if k in kw:
                kw[k] = tuple(kw[k])
        out.append(_CLS[name](**kw))
    return out


def readme_model():
    return (Lens(z=0.5, focal_length=1.0), Detector(z=1.0, pixel_size=(0.01, 0.01), shape=(128, 128)))


def two_beam_model(p):
    M1, F1, defocus = p["M1"], p["F1"], p["defocus"]
    L1_z1 = F1 * (1 / M1 - 1)
    L1_z2 = F1 * (1 - M1)
    src = ParallelBeam(z=0.0 - defocus, radius=1e-3)
    lens = Lens(z=abs(L1_z1), focal_length=F1)
    bip = Biprism(z=abs(L1_z2) + L1_z2 / 2, def_x=p["def_x"])
    det = Detector(z=abs(L1_z2) + L1_z2, pixel_size=(p["pixel_size"], p["pixel_size"]), shape=(1024, 1024))
    return [src, lens, bip, det]


def biprism_model(p, shape=(1024, 1024), pixel=55e-6 / 2):
    M1, F1, M2, F2, defocus = p["M1"], p["F1"], p["M2"], p["F2"], p["defocus"]
    L1_z1 = F1 * (1 / M1 - 1)
    L1_z2 = F1 * (1 - M1)
    L2_z1 = F2 * (1 / M2 - 1)
    L2_z2 = F2 * (1 - M2)
    L1_z1, L1_z2, L2_z1, L2_z2 = np.abs([L1_z1, L1_z2, L2_z1, L2_z2])
    src = ParallelBeam(z=0.0 + defocus, radius=p["aperture_radius"])
    lens1 = Lens(focal_length=F1, z=float(L1_z1))
    bip = Biprism(z=float(L1_z1 + L1_z2 / 2), rotation=0.0, def_x=p["def_x"])
    lens2 = Lens(focal_length=F2, z=float(L1_z1 + L1_z2 + L2_z1))
    det = Detector(z=float(L1_z1 + L1_z2 + L2_z1 + L2_z2), pixel_size=(pixel, pixel), shape=shape)
    return [src, lens1, bip, lens2, det]


def biprism_lens_prop_model(p):
    M1, F1, defocus, deflection = p["M1"], p["F1"], p["defocus"], p["deflection"]
    L1_z1 = F1 * (1 / M1 - 1)
    L1_z2 = F1 * (1 - M1)
    src = ParallelBeam(z=0.0 + defocus, radius=0.0)
    lens = Lens(focal_length=F1, z=abs(L1_z1))
    bip = Biprism(z=abs(L1_z1) + abs(L1_z2) / 2, rotation=0.0, def_x=deflection)
    det = Detector(z=abs(L1_z1) + abs(L1_z2), pixel_size=(0.01, 0.01), shape=(128, 128))
    return [src, lens, bip, det]


def six_component_column():
    """BASELINE config C4 column."""
    return [
        ParallelBeam(z=0.0, radius=0.2e-9),
        AberratedLensKrivanek(z=2.5e-4, focal_length=2.5e-4,
                              coeffs=KrivanekCoeffs(C30=1e8, C12=1e-1, C21=1e3, C23=1e3)),
        Deflector(z=3e-4, def_x=1e-6, def_y=-1e-6),
        Lens(z=3.5e-4, focal_length=1e-3),
        Biprism(z=4e-4, def_x=1e-6),
        Detector(z=5e-4, pixel_size=(1e-9, 1e-9), shape=(128, 128)),
    ]


def kitchen_sink_model():
    """Every opcode once, with a distance-0 step and a ThickLens z jump."""
    return [
        PointSource(z=-0.1, semi_conv=0.01),
        Scanner(z=0.0, scan_pos_x=0.013, scan_pos_y=-0.02, scan_tilt_x=1e-3, scan_tilt_y=-2e-3),
        Lens(z=0.2, focal_length=0.35),
        Plane(z=0.2),
        Deflector(z=0.3, def_x=2e-3, def_y=-1e-3),
        Rotator(z=0.35, angle=33.0),
        ThickLens(z_po=0.4, z_pi=0.47, focal_length=-0.8),
        Biprism(z=0.6, def_x=-3e-3),
        Descanner(z=0.7, scan_pos_x=0.013, scan_pos_y=-0.02, scan_tilt_x=1e-3, scan_tilt_y=-2e-3,
                  descan_error=DescanError(*np.linspace(-0.3, 0.4, 12))),
        ScanGrid(z=0.8, pixel_size=(1e-3, 1e-3), shape=(16, 16), rotation=13.0),
        Detector(z=1.0, pixel_size=(0.01, 0.01), shape=(64, 64), flip_y=True),
    ]


def random_rays(n, rng=None, scale=0.64, slope=0.5, z=0.0, pl=0.0):
    rng = rng or np.random.default_rng(SEED)
    return Ray(
        x=rng.uniform(-scale, scale, n), y=rng.uniform(-scale, scale, n),
        dx=rng.uniform(-slope, slope, n), dy=rng.uniform(-slope, slope, n),
        z=np.full(n, z), pathlength=np.full(n, pl), _one=np.ones(n),
    )


def gaussian_rays(x, y, *, dx=None, dy=None, z=0.0, wavelength=2e-12, w0=1e-9, amplitude=1.0,
                  theta=None, waist_xy=None, radii=None, pathlength=None):
    n = len(x)
    zeros = np.zeros(n)
    return GaussianRay(
        x=np.asarray(x, float), y=np.asarray(y, float),
        dx=zeros if dx is None else np.asarray(dx, float),
        dy=zeros if dy is None else np.asarray(dy, float),
        z=np.full(n, z), pathlength=zeros if pathlength is None else np.asarray(pathlength, float),
        _one=np.ones(n), amplitude=np.broadcast_to(np.asarray(amplitude, float), (n,)).copy(),
        waist_xy=np.full((n, 2), w0) if waist_xy is None else np.asarray(waist_xy, float),
        radii_of_curv=np.full((n, 2), np.inf) if radii is None else np.asarray(radii, float),
        wavelength=np.full(n, wavelength), theta=zeros if theta is None else np.asarray(theta, float),
    )


def aperture_diffraction_case(nb, shape):
    """BASELINE config C2 (examples/aperture_diffraction.ipynb cells 2,4,11,15) at (nb, shape)."""
    wavelength, w0, aperture_radius, F1 = 2e-12, 1e-9, 1e-7, 1e-2
    x, y = fibonacci_spiral(nb, aperture_radius, alpha=0)
    area = np.pi * aperture_radius ** 2
    amp = area / (w0 ** 2 * nb * np.pi)
    pixel = wavelength * F1 / 1e-6 * (512 / shape[0]) if shape[0] < 512 else wavelength * F1 / 1e-6
    model = [ParallelBeam(z=0.0, radius=aperture_radius), Lens(focal_length=F1, z=F1),
             Detector(z=2 * F1, pixel_size=(pixel, pixel), shape=shape)]
    return gaussian_rays(x, y, wavelength=wavelength, w0=w0, amplitude=amp), model


def biprism_case(nb, shape, fov=1024 * 55e-6 / 2, general=False, rng=None):
    """BASELINE config C3 (examples/biprism.ipynb cells 1,2,6) at (nb, shape); ``general``
    gives the non-separable variant of SURVEY.md section 8d."""
    p = dict(M1=-200, F1=0.0025, M2=-1500, F2=0.02, defocus=1e-9, def_x=-2e-5, aperture_radius=50e-9)
    wavelength, w0 = 2e-12, 1e-9
    x, y = fibonacci_spiral(nb, p["aperture_radius"], alpha=0)
    area = np.pi * p["aperture_radius"] ** 2
    amp = area / (w0 ** 2 * nb * np.pi)
    model = biprism_model(p, shape=shape, pixel=fov / shape[0])
    kw = {}
    if general:
        rng = rng or np.random.default_rng(SEED)
        kw = dict(theta=rng.uniform(-np.pi / 2, np.pi / 2, nb),
                  waist_xy=rng.uniform(0.5, 2.0, (nb, 2)) * w0)
        det = model[-1]
        model[-1] = Detector(z=det.z, pixel_size=det.pixel_size, shape=det.shape, rotation=17.0)
    g = gaussian_rays(x, y, z=model[0].z, wavelength=wavelength, w0=w0, amplitude=amp, **kw)
    return g, model


def stem4d_case(scan_shape=(16, 12), det_shape=(24, 20), rng=None, z_src=-1e-5):
    """BASELINE config C5 geometry at a reduced size: point source just above
    the sample (ScanGrid) plane, Scanner, Descanner with a random DescanError, Detector."""
    rng = rng or np.random.default_rng(SEED)
    scan_grid = ScanGrid(z=0.0, pixel_size=(1e-9, 1e-9), shape=scan_shape, rotation=13.0)
    detector = Detector(z=0.5, pixel_size=(55e-6, 55e-6), shape=det_shape, flip_y=True)
    err = DescanError(*rng.uniform(-1e-3, 1e-3, 12))
    src = PointSource(z=z_src, semi_conv=1e-2)

    def model_fn(spx, spy):
        return [src, scan_grid, Scanner(z=0.0, scan_pos_x=spx, scan_pos_y=spy),
                Descanner(z=0.1, scan_pos_x=spx, scan_pos_y=spy, descan_error=err), detector]
    return model_fn, scan_grid, detector
