"""Times make_gaussian_image_host (C2) end to end and prints the library's per-phase breakdown
(TG_HOST_TIMING=1)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dataclasses import fields, replace
from tests import models as M
from temgymcore_b200.gaussian import make_gaussian_image_host
g, model = M.aperture_diffraction_case(10000, (1024, 1024))
gp = replace(g, **{f.name: torch.as_tensor(getattr(g, f.name)).pin_memory() for f in fields(g)})
for _ in range(3):
    make_gaussian_image_host(gp, model, cull_bits=0, device=0)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    t0 = time.perf_counter()
    make_gaussian_image_host(gp, model, cull_bits=0, device=0)
    ts.append((time.perf_counter() - t0) * 1e3)
print("wall ms per call: median %.3f min %.3f" % (np.median(ts), min(ts)))
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    make_gaussian_image_host(gp, model, cull_bits=0, device=0)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
