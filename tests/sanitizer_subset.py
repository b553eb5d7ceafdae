"""Small invocations of the kernels that synchronise through shared memory, mbarriers, TMEM or cross-GPU flag
words, sized to finish under `compute-sanitizer --tool racecheck|synccheck|memcheck` in seconds.  Run by
tests/test_gpu_sanitizer.py (and by hand: `compute-sanitizer --tool racecheck python tests/sanitizer_subset.py all`).
Every section checks its result, so a sanitizer run is also a correctness run."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import temgym_oracle as O  # noqa: E402
from tests import models as M  # noqa: E402
from temgymcore_b200 import _lib as L  # noqa: E402


def rel_l2(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))


def to_cuda(g):
    from dataclasses import fields, replace
    return replace(g, **{f.name: torch.as_tensor(np.asarray(getattr(g, f.name), dtype=np.float64), device="cuda")
                         for f in fields(g)})


def sec_gemm():
    lib = L.load()
    gen = torch.Generator(device="cuda").manual_seed(3)
    for (m, n, k) in ((128, 256, 4096), (200, 136, 1000)):      # stream-K split / ragged single tiles
        A = torch.rand((m, k), generator=gen, device="cuda") * 2 - 1
        B = torch.rand((n, k), generator=gen, device="cuda") * 2 - 1
        Ah, Bh = A.half(), B.half()
        Al, Bl = (A - Ah.float()).half(), (B - Bh.float()).half()
        D = torch.empty((m, n), dtype=torch.float64, device="cuda")
        L.check(lib.tg_gemm_f16x3(m, n, k, Ah.data_ptr(), Al.data_ptr(), Bh.data_ptr(), Bl.data_ptr(), k,
                                  D.data_ptr(), n, 0, torch.cuda.current_stream().cuda_stream), "tg_gemm_f16x3")
        ref = A.double() @ B.double().T
        assert float((D - ref).norm() / ref.norm()) < 3e-6


def sec_field():
    from temgymcore_b200.gaussian import make_gaussian_image_device
    g, model = M.aperture_diffraction_case(300, (64, 256))
    ref = O.make_gaussian_image(g, model)
    for kw in (dict(method="sfu", cull_bits=0), dict(method="tensor"), dict(method="auto")):
        assert rel_l2(make_gaussian_image_device(to_cuda(g), model, **kw).cpu().numpy(), ref) < 1e-5, kw
    g, model = M.biprism_case(400, (96, 256), fov=3 * 1024 * 55e-6 / 2)          # narrow beamlets: gather mode
    ref = O.make_gaussian_image(g, model)
    assert rel_l2(make_gaussian_image_device(to_cuda(g), model, method="sfu").cpu().numpy(), ref) < 1e-5


def sec_binned():
    # tile-binned sum: atomicOr bitmaps, ordered expansion, ragged stream-K schedule (partial tiles through scratch slots
    # + arrival counters), tiled operands; 900 narrow beamlets on 3 x 7 tiles, and a row block
    from temgymcore_b200.gaussian import _field_sum_grid, beamlet_polynomials
    g, model = M.biprism_case(900, (300, 416), fov=3 * 1024 * 55e-6 / 2)
    poly, n, dev = beamlet_polynomials(to_cuda(g), model)
    dense = _field_sum_grid(poly, n, model[-1], dev, cull_bits=0, method="sfu").cpu().numpy()
    got = _field_sum_grid(poly, n, model[-1], dev, cull_bits=40, method="tensor_binned").cpu().numpy()
    assert rel_l2(got, dense) < 3e-6
    rows = _field_sum_grid(poly, n, model[-1], dev, cull_bits=40, method="tensor_binned", row0=70, nrows=200).cpu().numpy()
    assert rel_l2(rows, dense[70:270]) < 3e-6


def sec_stem4d():
    from temgymcore_b200.stem4d import backproject_4dstem, system_geometry
    fn, sg, det = M.stem4d_case((8, 6), (64, 48))
    geo = system_geometry(fn, sg, det)
    data = torch.randint(0, 5, (48, 64, 48), device="cuda").float()
    outs = []
    for kern in ("stepwise", "auto"):
        img = torch.zeros(tuple(sg.shape), dtype=torch.float32, device="cuda")
        backproject_4dstem(data, None, sg, det, scan_range=(0, 48), out=img, geometry=geo, kernel=kern)
        outs.append(img.cpu().numpy())
    np.testing.assert_array_equal(outs[0], outs[1])       # integer data: exact


def sec_trace():
    from temgymcore_b200.ray import RAY_FIELDS, Ray
    from temgymcore_b200.run import run_to_end_abcd
    rays = M.random_rays(1000)
    dr = Ray(*(torch.as_tensor(getattr(rays, f), device="cuda") for f in RAY_FIELDS))
    out, abcd = run_to_end_abcd(dr, M.kitchen_sink_model())       # AoS output staged in smem + bulk store
    _, ref = O.abcd_run_to_end(rays, M.kitchen_sink_model())
    np.testing.assert_allclose(abcd.cpu().numpy(), ref, rtol=1e-12, atol=1e-15)


def sec_jets():
    # barriers inside the Krivanek lens (per-ray table of partials in smem), smem tensor assembly, bulk stores;
    # 203 rays: a ragged last CTA at every order
    from temgymcore_b200.ray import RAY_FIELDS, Ray
    from temgymcore_b200.run import calculate_derivatives
    rays = M.random_rays(203, scale=0.2e-9, slope=1e-6)
    model = M.six_component_column()
    dr = Ray(*(torch.as_tensor(getattr(rays, f), device="cuda") for f in RAY_FIELDS))
    ref = O.calculate_derivatives(rays, model, 3)
    for order in (1, 2, 3):
        got = calculate_derivatives(dr, model, order)
        for k in range(order):
            g = got[k].tensor.cpu().numpy()
            for f in range(7):
                np.testing.assert_allclose(g[:, f], ref[k][:, f], rtol=1e-8, atol=1e-8 * np.abs(ref[k][:, f]).max())


def sec_peer():
    from temgymcore_b200.distributed import PeerImage
    from temgymcore_b200.gaussian import beamlet_polynomials
    g, model = M.aperture_diffraction_case(200, (64, 128))
    poly, nb, dev = beamlet_polynomials(to_cuda(g), model)
    with PeerImage(64, 128) as pimg:                               # world of one rank: flags + barrier kernel
        for _ in range(2):
            pimg.field_sum(poly, nb, model[-1], cull_bits=0, method="auto")
            pimg.barrier()
        assert rel_l2(pimg.image.cpu().numpy(), O.make_gaussian_image(g, model)) < 1e-5


SECTIONS = {"gemm": sec_gemm, "field": sec_field, "binned": sec_binned, "stem4d": sec_stem4d, "trace": sec_trace, "jets": sec_jets,
            "peer": sec_peer}

if __name__ == "__main__":
    which = sys.argv[1:] or ["all"]
    names = list(SECTIONS) if "all" in which else which
    torch.cuda.set_device(0)
    for n in names:
        SECTIONS[n]()
        torch.cuda.synchronize()
        print(f"section {n}: ok", flush=True)
