"""C3 (1e5 beamlets, 2048^2, 40-bit culling): the tile-binned tensor-core sum against the culled SFU kernel and the dense
GEMM -- per-call times (eager and as a graph plan) and the images' distance.  `python tools/exp_binned.py [nb] [size]`"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dataclasses import fields, replace
from tests import models as M
from temgymcore_b200.gaussian import GaussianImagePlan, make_gaussian_image_device
quick = "quick" in sys.argv            # the tile-binned plan only (A/B runs of its knobs)
args = [a for a in sys.argv[1:] if a != "quick"]
nb = int(args[0]) if len(args) > 0 else 100_000
size = int(args[1]) if len(args) > 1 else 2048
dev = torch.device("cuda", 0)
g3, model3 = M.biprism_case(nb, (size, size))
g3d = replace(g3, **{f.name: torch.as_tensor(getattr(g3, f.name), device=dev) for f in fields(g3)})
imgs = {}
for method, cull in (() if quick else (("sfu", 40), ("tensor_binned", 40), ("auto", 40), ("tensor_binned", 24), ("tensor", 0))):
    ts = []
    for _ in range(5):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        img = make_gaussian_image_device(g3d, model3, cull_bits=cull, method=method)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    imgs[(method, cull)] = img
    print(f"eager {method:14s} cull {cull:2d}: " + " ".join(f"{t:.3f}" for t in ts) + " ms", flush=True)
ref = imgs.get(("tensor", 0))
for k, v in imgs.items():
    print(f"rel L2 of {k} against the dense GEMM: {float((v - ref).norm() / ref.norm()):.3e}", flush=True)
del imgs
for method, cull in ((("tensor_binned", 40),) if quick else (("tensor_binned", 40), ("sfu", 40), ("auto", 40))):
    plan = GaussianImagePlan(g3d, model3, cull_bits=cull, method=method)
    for _ in range(3):
        plan.run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"plan  {method:14s} ({plan.method}) cull {cull:2d}: median {np.median(ts):.3f} min {min(ts):.3f} ms", flush=True)
    del plan
