"""Launches exactly the kernels we profile, a few times each, so `ncu -k regex:... -s N -c 1`
hits a warm, full-size launch.  Usage: python tools/prof_kernels.py [field|gemm|trace|trace_c4|stem4d|jets|field_c3|all]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import models as M  # noqa: E402
from temgymcore_b200.gaussian import _field_sum_grid, beamlet_polynomials  # noqa: E402
from temgymcore_b200.ray import RAY_FIELDS, Ray  # noqa: E402
from temgymcore_b200.run import run_to_end_abcd  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda", 0)
if what in ("field", "all"):
    g, model = M.aperture_diffraction_case(10_000, (1024, 1024))
    poly, n, _ = beamlet_polynomials(g, model)
    for _ in range(3):
        _field_sum_grid(poly, n, model[-1], dev, cull_bits=0)
    torch.cuda.synchronize()
if what in ("gemm",):
    g, model = M.aperture_diffraction_case(10_000, (1024, 1024))
    poly, n, _ = beamlet_polynomials(g, model)
    for _ in range(3):
        _field_sum_grid(poly, n, model[-1], dev, method="tensor")
    torch.cuda.synchronize()
if what in ("trace", "all"):
    rr = M.random_rays(10_000_000)
    rd = Ray(*(torch.as_tensor(getattr(rr, f), device=dev) for f in RAY_FIELDS))
    for _ in range(3):
        o = run_to_end_abcd(rd, M.readme_model())
    torch.cuda.synchronize()
if what in ("trace_c4", "all"):
    rr = M.random_rays(10_000_000, scale=0.2e-9, slope=1e-6)
    rd = Ray(*(torch.as_tensor(getattr(rr, f), device=dev) for f in RAY_FIELDS))
    for _ in range(3):
        o = run_to_end_abcd(rd, M.six_component_column())
    torch.cuda.synchronize()
if what in ("stem4d", "all"):
    from temgymcore_b200.stem4d import backproject_4dstem, system_geometry
    fn, sg, det = M.stem4d_case((256, 256), (256, 256), z_src=-1e-6)
    geo = system_geometry(fn, sg, det)
    data = torch.rand((65536, 256, 256), device=dev, dtype=torch.float32)
    img = torch.zeros((256, 256), dtype=torch.float32, device=dev)
    for _ in range(3):
        backproject_4dstem(data, None, sg, det, scan_range=(0, 65536), out=img, geometry=geo)
    torch.cuda.synchronize()
if what in ("jets", "all"):
    from temgymcore_b200.run import calculate_derivatives
    rr = M.random_rays(200_000)
    rd = Ray(*(torch.as_tensor(getattr(rr, f), device=dev) for f in RAY_FIELDS))
    for _ in range(3):
        d = calculate_derivatives(rd, M.readme_model(), 3)
    torch.cuda.synchronize()
if what in ("jets_c4",):
    from temgymcore_b200.run import calculate_derivatives
    rr = M.random_rays(200_000, scale=0.2e-9, slope=1e-9)
    rd = Ray(*(torch.as_tensor(getattr(rr, f), device=dev) for f in RAY_FIELDS))
    for _ in range(3):
        d = calculate_derivatives(rd, M.six_component_column(), 3)
    torch.cuda.synchronize()
if what in ("c3_tensor",):
    from dataclasses import fields, replace
    from temgymcore_b200.gaussian import make_gaussian_image_device
    g3, model3 = M.biprism_case(100_000, (2048, 2048))
    g3d = replace(g3, **{f.name: torch.as_tensor(getattr(g3, f.name), device=dev) for f in fields(g3)})
    for _ in range(2):
        make_gaussian_image_device(g3d, model3, cull_bits=0, method="tensor")
    torch.cuda.synchronize()
if what in ("field_c3",):
    from dataclasses import fields, replace
    from temgymcore_b200.gaussian import make_gaussian_image_device
    g3, model3 = M.biprism_case(100_000, (2048, 2048))
    g3d = replace(g3, **{f.name: torch.as_tensor(getattr(g3, f.name), device=dev) for f in fields(g3)})
    for _ in range(3):
        make_gaussian_image_device(g3d, model3, method="sfu")
    torch.cuda.synchronize()
if what in ("gemm_shard",):
    # the stream-K shape of one rank's 128-row shard / a 256-row block of the host pipeline (C2 beamlets)
    g, model = M.aperture_diffraction_case(10_000, (1024, 1024))
    poly, n, _ = beamlet_polynomials(g, model)
    for nrows in (128, 256):
        for _ in range(3):
            _field_sum_grid(poly, n, model[-1], dev, row0=256, nrows=nrows, method="tensor")
    torch.cuda.synchronize()
