#!/usr/bin/env bash
mkdir -p gpurun_out
echo "== pair kernel: small sanity first (timeout guards a hang)"
TG_GEMM_PAIR=1 timeout 120 python - > gpurun_out/j6_sanity.log 2>&1 <<'PY'
import torch, sys, os
sys.path.insert(0, os.getcwd())
from temgymcore_b200 import _lib as L
lib = L.load()
gen = torch.Generator(device="cuda").manual_seed(1)
for (m, n, k) in ((256, 128, 64), (256, 128, 512), (256, 256, 4096), (512, 384, 1000), (300, 200, 333), (1024, 2048, 2000)):
    A = torch.rand((m, k), generator=gen, device="cuda") * 2 - 1
    B = torch.rand((n, k), generator=gen, device="cuda") * 2 - 1
    ldk = ((k + 7) // 8) * 8
    Ap = torch.zeros((m, ldk), device="cuda"); Ap[:, :k] = A
    Bp = torch.zeros((n, ldk), device="cuda"); Bp[:, :k] = B
    Ah, Bh = Ap.half(), Bp.half()
    Al, Bl = (Ap - Ah.float()).half(), (Bp - Bh.float()).half()
    D = torch.full((m, n), -7.0, dtype=torch.float64, device="cuda")
    L.check(lib.tg_gemm_f16x3(m, n, k, Ah.data_ptr(), Al.data_ptr(), Bh.data_ptr(), Bl.data_ptr(), ldk,
                              D.data_ptr(), n, 0, torch.cuda.current_stream().cuda_stream), "gemm")
    torch.cuda.synchronize()
    ref = A.double() @ B.double().T
    print((m, n, k), "rel err", float((D - ref).norm() / ref.norm()), flush=True)
print("SANITY DONE")
PY
cat gpurun_out/j6_sanity.log | tail -12
if grep -q "SANITY DONE" gpurun_out/j6_sanity.log; then
  echo "== pair kernel tests"
  TG_GEMM_PAIR=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q --timeout 300 -k "test_gemm_f16x3 or test_gemm_tf32x3 or tensor_path_parity or c2_full_size_tensor" > gpurun_out/j6_pytest.log 2>&1
  tail -6 gpurun_out/j6_pytest.log
  echo "== timings"
  TG_GEMM_PAIR=1 timeout 300 python tools/exp_gemm.py 2>&1 | tail -5
  TG_GEMM_PAIR=0 timeout 300 python tools/exp_gemm.py 2>&1 | tail -5
  for p in 1 0; do TG_GEMM_PAIR=$p timeout 600 python bench.py --steps 10 --warmup 3 --skip-c3 --no-cpu-baseline 2>/dev/null | grep -E '"section": "(headline|e2e|roofline_tensor_path)"' | cut -c1-700; done
fi
