"""C3 (1e5 beamlets, 2048^2) per-call times of the field-sum methods."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dataclasses import fields, replace
from tests import models as M
from temgymcore_b200.gaussian import make_gaussian_image_device
dev = torch.device("cuda", 0)
g3, model3 = M.biprism_case(100_000, (2048, 2048))
g3d = replace(g3, **{f.name: torch.as_tensor(getattr(g3, f.name), device=dev) for f in fields(g3)})
for method, cull in (("tensor", 0), ("auto", 0), ("tensor", 0), ("auto", 40), ("sfu", 40), ("tensor_tf32", 0)):
    ts = []
    for _ in range(5):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        make_gaussian_image_device(g3d, model3, cull_bits=cull, method=method)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(method, cull, " ".join(f"{t:.2f}" for t in ts), flush=True)
