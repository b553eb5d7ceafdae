"""Fixed cost of a split-K GEMM launch: the 4-multiplication real GEMM (tg_gemm_f16x3) on row blocks of 128 / 256 /
512 rows for several depths K -- the intercept of t(K) is what a launch costs beyond its tensor work."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from temgymcore_b200 import _lib as L
lib = L.load()
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
gen = torch.Generator(device=dev).manual_seed(1)
st = torch.cuda.current_stream().cuda_stream
N = 2048
def timeit(run, flushing=True):
    for _ in range(3): run()
    ts = []
    for _ in range(10):
        if flushing: flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
for nb in (1248, 2496, 5000, 10000, 20000):
    K = 2 * nb
    B = (torch.rand((N, K), generator=gen, device=dev) * 2 - 1)
    Bh = B.half(); Bl = (B - Bh.float()).half(); del B
    for M in (128, 256, 512):
        A = (torch.rand((M, K), generator=gen, device=dev) * 2 - 1)
        Ah = A.half(); Al = (A - Ah.float()).half(); del A
        D = torch.empty((M, N), dtype=torch.float64, device=dev)
        def run():
            L.check(lib.tg_gemm_f16x3(M, N, K, Ah.data_ptr(), Al.data_ptr(), Bh.data_ptr(), Bl.data_ptr(), K, D.data_ptr(), N, 0, st), "gemm")
        cold, warm = timeit(run, True), timeit(run, False)
        sc = (C := __import__("ctypes")).c_int32 * 10
        s = sc(); lib.tg_gemm_schedule(M, N, K, 1, 148, -1, None, 0, s)
        print(f"K={K:6d} M={M:4d}: cold {cold*1e3:7.1f} us  warm(L2) {warm*1e3:7.1f} us   G={s[4]} parts={s[9]} q={s[6]} nch={s[3]}", flush=True)
    del Bh, Bl
