"""Is the fp16x3 GEMM clock / power limited?  Runs the C2-shaped GEMM back to back for a few seconds while
nvidia-smi samples SM clock and power, and compares the sustained time per launch with the isolated one."""
import os, subprocess, sys, threading, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from temgymcore_b200 import _lib as L
lib = L.load()
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(1)
M_, N, K = 1024, 2048, 20000
A32 = torch.rand((M_, K), generator=gen, device=dev) * 2 - 1
B32 = torch.rand((N, K), generator=gen, device=dev) * 2 - 1
Ah = A32.half(); Al = (A32 - Ah.float()).half(); Bh = B32.half(); Bl = (B32 - Bh.float()).half()
D = torch.empty((M_, N), dtype=torch.float64, device=dev)
def run():
    L.check(lib.tg_gemm_f16x3(M_, N, K, Ah.data_ptr(), Al.data_ptr(), Bh.data_ptr(), Bl.data_ptr(), K,
                              D.data_ptr(), N, 0, torch.cuda.current_stream().cuda_stream), "gemm")
rows = []
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap",
                         "--format=csv,noheader,nounits", "-lms", "20", "-i", "0"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [rows.append((time.time(), l.strip())) for l in proc.stdout], daemon=True).start()
for _ in range(5): run()
torch.cuda.synchronize()
time.sleep(0.5)
t_idle_end = time.time()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 8000
e0.record()
for _ in range(n): run()
e1.record(); torch.cuda.synchronize()
t_load_end = time.time()
ms = e0.elapsed_time(e1) / n
time.sleep(0.3); proc.terminate()
load = [r for t, r in rows if t_idle_end + 0.3 < t < t_load_end]
clk = [float(r.split(",")[0]) for r in load]; pw = [float(r.split(",")[1]) for r in load]
cap = sum("Active" == r.split(",")[2].strip() for r in load)
print(f"sustained: {ms:.4f} ms per GEMM = {3*2.0*M_*N*K/(ms*1e-3)/1e12:.0f} TF/s executed; under load: SM clock median "
      f"{np.median(clk):.0f} MHz (min {min(clk):.0f}), power median {np.median(pw):.0f} W (max {max(pw):.0f}), "
      f"sw_power_cap active in {cap}/{len(load)} samples")
# torch bf16 matmul for comparison, same protocol
a = torch.randn((8192, 8192), device=dev, dtype=torch.bfloat16); b = torch.randn((8192, 8192), device=dev, dtype=torch.bfloat16)
for _ in range(3): a @ b
torch.cuda.synchronize()
e0.record()
for _ in range(50): a @ b
e1.record(); torch.cuda.synchronize()
print(f"cuBLAS bf16 8192^3 sustained 50 launches: {2*8192**3/(e0.elapsed_time(e1)/50*1e-3)/1e12:.0f} TF/s")
