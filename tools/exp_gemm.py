"""GEMM experiments: fp16x3 kernel on the C2 shape and on row shards / row blocks, CUDA events, L2 flushed.
   TG_GEMM_STREAMK=0|1 python tools/exp_gemm.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from temgymcore_b200 import _lib as L
lib = L.load()
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
gen = torch.Generator(device=dev).manual_seed(1)
N, K = 2048, 20000
B32 = torch.rand((N, K), generator=gen, device=dev) * 2 - 1
Bh = B32.half(); Bl = (B32 - Bh.float()).half()
for Mrows in (1024, 512, 256, 128):
    A32 = torch.rand((Mrows, K), generator=gen, device=dev) * 2 - 1
    Ah = A32.half(); Al = (A32 - Ah.float()).half()
    D = torch.empty((Mrows, N), dtype=torch.float64, device=dev)
    def run():
        L.check(lib.tg_gemm_f16x3(Mrows, N, K, Ah.data_ptr(), Al.data_ptr(), Bh.data_ptr(), Bl.data_ptr(), K,
                                  D.data_ptr(), N, 0, torch.cuda.current_stream().cuda_stream), "gemm")
    for _ in range(3): run()
    ts = []
    for _ in range(10):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    ref = A32[:64].double() @ B32.double().T
    err = float((D[:64] - ref).norm() / ref.norm())
    print(f"STREAMK={os.environ.get('TG_GEMM_STREAMK','1')} M={Mrows}: {ms:.4f} ms (min {min(ts):.4f}) "
          f"executed {3*2.0*Mrows*N*K/(ms*1e-3)/1e12:.0f} TF/s, rel err {err:.2e}", flush=True)
