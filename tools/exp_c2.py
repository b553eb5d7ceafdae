"""C2 headline step (GaussianImagePlan replay, L2 flushed between replays) and its distance from the SFU image:
TG_GEMM_TILED=0/1 python tools/exp_c2.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dataclasses import fields, replace
from tests import models as M
from temgymcore_b200.gaussian import GaussianImagePlan, make_gaussian_image_device
dev = torch.device("cuda", 0)
g, model = M.aperture_diffraction_case(10_000, (1024, 1024))
gd = replace(g, **{f.name: torch.as_tensor(getattr(g, f.name), device=dev) for f in fields(g)})
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
plan = GaussianImagePlan(gd, model, cull_bits=0)
for _ in range(5):
    plan.run()
ts = []
for _ in range(30):
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = plan.run(); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ref = make_gaussian_image_device(gd, model, cull_bits=0, method="sfu")
print(f"TILED={os.environ.get('TG_GEMM_TILED', '1')} C2 plan ({plan.method}): median {np.median(ts):.4f} ms min {min(ts):.4f} ms, "
      f"rel L2 vs SFU {float((out - ref).norm() / ref.norm()):.2e}", flush=True)
