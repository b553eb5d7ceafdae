"""GEMM experiments, round 2: the 4-multiplication real GEMM (tg_gemm_f16x3, K = 2 nb) against the 3-product
complex GEMM (tg_cgemm3_f16x3, K'' = 3 nb) on the C2 shape and on row shards; CUDA events, L2 flushed.
   [TG_LIB_PATH=...ck256.so] [TG_GEMM_STREAMK=0|1|2] python tools/exp_gemm2.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from temgymcore_b200 import _lib as L
lib = L.load()
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
gen = torch.Generator(device=dev).manual_seed(1)
W, nb = 1024, 10000
kch = lib.tg_gemm_chunk_k()
tag = f"chunk={kch} STREAMK={os.environ.get('TG_GEMM_STREAMK', '1')}"

def split(x):
    hi = x.half()
    return hi, (x - hi.float()).half()

def pack3(Xr, Xi, which):
    rows, n = Xr.shape
    g = (n + kch - 1) // kch
    pad = g * kch - n
    if pad:
        z = torch.zeros((rows, pad), dtype=Xr.dtype, device=Xr.device)
        Xr, Xi = torch.cat([Xr, z], 1), torch.cat([Xi, z], 1)
    blocks = (Xr + Xi, Xr, Xi) if which == "A" else (Xr, Xi - Xr, Xr + Xi)
    return torch.stack([b.reshape(rows, g, kch) for b in blocks], dim=2).reshape(rows, g * 3 * kch).contiguous()

def timeit(run):
    for _ in range(3): run()
    ts = []
    for _ in range(10):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(min(ts))

Vr, Vi = (torch.rand((W, nb), generator=gen, device=dev) * 2 - 1 for _ in range(2))
# real formulation operands: B'[2c] = (Vr, -Vi) interleaved, B'[2c+1] = (Vi, Vr)
Breal = torch.empty((2 * W, 2 * nb), device=dev)
Breal[0::2, 0::2], Breal[0::2, 1::2], Breal[1::2, 0::2], Breal[1::2, 1::2] = Vr, -Vi, Vi, Vr
Brh, Brl = split(Breal)
B3h, B3l = split(pack3(Vr, Vi, "B"))
del Breal
st = torch.cuda.current_stream().cuda_stream
for Mrows in (1024, 512, 256, 128):
    Ur, Ui = (torch.rand((Mrows, nb), generator=gen, device=dev) * 2 - 1 for _ in range(2))
    Areal = torch.empty((Mrows, 2 * nb), device=dev)
    Areal[:, 0::2], Areal[:, 1::2] = Ur, Ui
    Arh, Arl = split(Areal)
    A3h, A3l = split(pack3(Ur, Ui, "A"))
    K3 = A3h.shape[1]
    D = torch.empty((Mrows, 2 * W), dtype=torch.float64, device=dev)
    ref = (torch.complex(Ur[:64].double(), Ui[:64].double()) @ torch.complex(Vr.double(), Vi.double()).T)
    def run4():
        L.check(lib.tg_gemm_f16x3(Mrows, 2 * W, 2 * nb, Arh.data_ptr(), Arl.data_ptr(), Brh.data_ptr(), Brl.data_ptr(),
                                  2 * nb, D.data_ptr(), 2 * W, 0, st), "gemm4")
    def run3():
        L.check(lib.tg_cgemm3_f16x3(Mrows, W, K3, A3h.data_ptr(), A3l.data_ptr(), B3h.data_ptr(), B3l.data_ptr(), K3,
                                    D.data_ptr(), 2 * W, 0, st), "gemm3")
    for name, run, macs in (("4-mult real", run4, 4), ("3-product  ", run3, 3)):
        med, mn = timeit(run)
        got = torch.view_as_complex(D[:64].reshape(64, W, 2).contiguous())
        err = float((got - ref).norm() / ref.norm())
        print(f"{tag} M={Mrows:5d} {name}: {med:.4f} ms (min {mn:.4f})  executed {3*2.0*macs*Mrows*W*nb/(med*1e-3)/1e12:6.0f} TF/s"
              f"  algorithmic {8.0*Mrows*W*nb/(med*1e-3)/1e12:5.0f} TF/s  rel err {err:.2e}", flush=True)
