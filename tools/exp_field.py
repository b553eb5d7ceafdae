"""SFU field kernel: accuracy vs the fp64 oracle on a small dense case and timing on C2 (dense)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import temgym_oracle as O
from tests import models as M
from temgymcore_b200.gaussian import _field_sum_grid, beamlet_polynomials, make_gaussian_image
dev = torch.device("cuda", 0)
for name, (g, model) in (("c2 small", M.aperture_diffraction_case(2000, (256, 256))),
                         ("c3 general small", M.biprism_case(3000, (256, 256), general=True))):
    got = np.asarray(make_gaussian_image(g, model, cull_bits=0, method="sfu"))
    ref = O.make_gaussian_image(g, model)
    print(name, "rel L2 vs oracle", np.linalg.norm(got - ref) / np.linalg.norm(ref), flush=True)
g, model = M.aperture_diffraction_case(10_000, (1024, 1024))
poly, n, _ = beamlet_polynomials(g, model)
for _ in range(3):
    _field_sum_grid(poly, n, model[-1], dev, cull_bits=0, method="sfu")
torch.cuda.synchronize()
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); _field_sum_grid(poly, n, model[-1], dev, cull_bits=0, method="sfu"); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts))
print("C2 dense SFU path: %.3f ms -> %.3e evals/s" % (ms, n * 1024 * 1024 / ms * 1e3))
