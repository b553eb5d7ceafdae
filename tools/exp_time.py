"""Ad-hoc kernel timing for experiments (CUDA events, warm): python tools/exp_time.py [c4|stem4d|c1] ..."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import models as M  # noqa: E402
from temgymcore_b200.ray import RAY_FIELDS, Ray  # noqa: E402
from temgymcore_b200.run import RayTracePlan  # noqa: E402

dev = torch.device("cuda", 0)


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


for what in sys.argv[1:]:
    if what in ("c4", "c1"):
        n = 10_000_000
        rr = (M.random_rays(n, scale=0.2e-9, slope=1e-9) if what == "c4" else M.random_rays(n))
        rd = Ray(*(torch.as_tensor(getattr(rr, f), device=dev) for f in RAY_FIELDS))
        plan = RayTracePlan(rd, M.six_component_column() if what == "c4" else M.readme_model())
        med, mn = timed(plan.run)
        print(f"{what} 1e7 rays (env TG_KRIV_MINB={os.environ.get('TG_KRIV_MINB')}): median {med:.4f} ms  min {mn:.4f} ms "
              f"-> {n * 312 / (med * 1e-3) / 1e9:.0f} GB/s", flush=True)
        del rd, plan
    if what == "stem4d":
        from temgymcore_b200.stem4d import backproject_4dstem, system_geometry
        fn, sg, det = M.stem4d_case((256, 256), (256, 256), z_src=-1e-6)
        geo = system_geometry(fn, sg, det)
        data = torch.rand((65536, 256, 256), device=dev, dtype=torch.float32)
        img = torch.zeros((256, 256), dtype=torch.float32, device=dev)
        for k in ("auto", "dda", "affine"):
            med, mn = timed(lambda: backproject_4dstem(data, None, sg, det, scan_range=(0, 65536), out=img,
                                                       geometry=geo, kernel=k), n=5, warm=2)
            print(f"stem4d (TG_DDA_MINB={os.environ.get('TG_DDA_MINB')}) kernel={k}: median {med:.3f} ms min {mn:.3f} ms -> {17.18 / med:.2f} TB/s", flush=True)
        d16 = (data * 100).to(torch.uint16)
        med, mn = timed(lambda: backproject_4dstem(d16, None, sg, det, scan_range=(0, 65536), out=img, geometry=geo),
                        n=5, warm=2)
        print(f"stem4d uint16 auto: median {med:.3f} ms -> {8.59 / med:.2f} TB/s", flush=True)
        del data, d16
    torch.cuda.empty_cache()
