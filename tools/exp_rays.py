"""C1 / C4 ray + ABCD kernel at 1e6 and 1e7 rays as RayTracePlan replays with L2 flushed between replays:
TG_TRACE_PERSIST=<CTAs per SM> python tools/exp_rays.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import models as M
from temgymcore_b200.ray import RAY_FIELDS, Ray
from temgymcore_b200.run import RayTracePlan
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for what, n in (("c1", 1_000_000), ("c1", 10_000_000), ("c4", 10_000_000)):
    rr = M.random_rays(n, scale=0.2e-9, slope=1e-9) if what == "c4" else M.random_rays(n)
    rd = Ray(*(torch.as_tensor(getattr(rr, f), device=dev) for f in RAY_FIELDS))
    plan = RayTracePlan(rd, M.six_component_column() if what == "c4" else M.readme_model())
    for _ in range(3):
        plan.run()
    ts = []
    for _ in range(15):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plan.run(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    med = float(np.median(ts))
    print(f"PERSIST={os.environ.get('TG_TRACE_PERSIST', '0'):>2s} {what} {n:>9d} rays: median {med:.4f} ms min {min(ts):.4f} "
          f"-> {n * 312 / (med * 1e-3) / 1e9:.0f} GB/s", flush=True)
    del rd, plan
