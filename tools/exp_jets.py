"""Order-3 jets (calculate_derivatives) on the README model and the six-component Krivanek column: CUDA events."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import models as M
from temgymcore_b200.ray import RAY_FIELDS, Ray
from temgymcore_b200.run import calculate_derivatives
dev = torch.device("cuda", 0)
n = 200_000
BYTES = 56 + 8 * (49 + 343 + 2401)
for which in ("c1", "c4"):
    rng = np.random.default_rng(M.SEED)
    rr = M.random_rays(n, rng) if which == "c1" else M.random_rays(n, rng, scale=0.2e-9, slope=1e-9)
    model = M.readme_model() if which == "c1" else M.six_component_column()
    rd = Ray(*(torch.as_tensor(getattr(rr, f), device=dev) for f in RAY_FIELDS))
    for order in (1, 2, 3):
        ts = []
        for it in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); d = calculate_derivatives(rd, model, order); e1.record(); torch.cuda.synchronize()
            if it >= 2: ts.append(e0.elapsed_time(e1))
            del d
        ms = float(np.median(ts))
        b = 56 + 8 * sum((49, 343, 2401)[:order])
        print(f"{which} order {order}: {ms:.4f} ms per {n} rays, {n * b / ms / 1e6:.0f} GB/s", flush=True)
