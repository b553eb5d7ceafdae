"""Which path does method="auto" take (bitwise comparison with the sfu / tensor results), and how long do
the three take, for a few narrow-beamlet cases?  Run with TG_SFU_WINS_BELOW=<threshold>."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import models as M
from temgymcore_b200.gaussian import _field_sum_grid, beamlet_polynomials
dev = torch.device("cuda", 0)


def t(fn, n=3):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / n


print("threshold", os.environ.get("TG_SFU_WINS_BELOW", "default"))
for nb, shape in ((4000, (1024, 1024)), (4000, (2048, 2048)), (20000, (2048, 2048)), (100000, (2048, 2048))):
    g, model = M.biprism_case(nb, shape)
    poly, n, _ = beamlet_polynomials(g, model)
    a, ta = t(lambda: _field_sum_grid(poly, n, model[-1], dev, method="auto"))
    s_, ts = t(lambda: _field_sum_grid(poly, n, model[-1], dev, method="sfu"))
    te, tt = t(lambda: _field_sum_grid(poly, n, model[-1], dev, method="tensor"))
    path = "sfu" if torch.equal(a, s_) else ("tensor" if torch.equal(a, te) else "?")
    print(f"nb={nb} {shape}: auto -> {path} {ta:.3f} ms | sfu {ts:.3f} ms | tensor {tt:.3f} ms", flush=True)
