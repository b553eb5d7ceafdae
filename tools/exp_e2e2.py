"""End-to-end host-buffer C2 image: wall time per call, the library's phase breakdown (TG_HOST_TIMING), raw PCIe
   copy rates.  TG_E2E_BLOCK_ROWS=<rows|-1> python tools/exp_e2e2.py [packed]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dataclasses import fields, replace
from tests import models as M
from temgymcore_b200.gaussian import make_gaussian_image_host, pack_beamlets_pinned
g, model = M.aperture_diffraction_case(10000, (1024, 1024))
packed = "packed" in sys.argv
gp = pack_beamlets_pinned(g) if packed else replace(g, **{f.name: torch.as_tensor(getattr(g, f.name)).pin_memory() for f in fields(g)})
for _ in range(5):
    make_gaussian_image_host(gp, model, cull_bits=0, device=0)
torch.cuda.synchronize()
ts = []
for _ in range(20):
    t0 = time.perf_counter()
    make_gaussian_image_host(gp, model, cull_bits=0, device=0)
    ts.append((time.perf_counter() - t0) * 1e3)
print(f"BLOCK_ROWS={os.environ.get('TG_E2E_BLOCK_ROWS')} packed={packed}: wall ms per call median {np.median(ts):.3f} min {min(ts):.3f}", flush=True)
if "pcie" in sys.argv:
    d = torch.empty(16 << 20, dtype=torch.uint8, device="cuda")
    h = torch.empty(16 << 20, dtype=torch.uint8, pin_memory=True)
    for name, fn in (("D2H", lambda: h.copy_(d, non_blocking=True)), ("H2D", lambda: d.copy_(h, non_blocking=True))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        print(f"PCIe {name} 16 MiB pinned: {16.777216 * 10 / e0.elapsed_time(e1):.1f} GB/s", flush=True)
    hs = torch.empty(4 << 20, dtype=torch.uint8, pin_memory=True)
    ds = d[:4 << 20]
    for _ in range(3): hs.copy_(ds, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): hs.copy_(ds, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(f"PCIe D2H 4 MiB pinned: {4.194304 * 10 / e0.elapsed_time(e1):.1f} GB/s", flush=True)
