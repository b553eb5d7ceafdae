#!/usr/bin/env bash
mkdir -p gpurun_out
for pf in 0 4 6 8 12; do echo "prefetch=$pf"; TG_GEMM_PREFETCH=$pf timeout 200 python tools/exp_gemm.py 2>&1 | tail -4; done
echo "== correctness of the prefetching stream-K (default)"
timeout 600 python -m pytest tests/test_gpu_parity.py -q --timeout 300 -k "test_gemm_f16x3 or streamk or cgemm3 or host_pipeline" 2>&1 | tail -3
echo "== e2e"
for br in 256 512; do TG_E2E_BLOCK_ROWS=$br timeout 200 python tools/exp_e2e2.py packed 2>&1 | tail -1; done
TG_E2E_BLOCK_ROWS=256 TG_HOST_TIMING=1 timeout 200 python tools/exp_e2e2.py packed 2>&1 | tail -3
