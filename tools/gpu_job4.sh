#!/usr/bin/env bash
mkdir -p gpurun_out
echo "== sanity (3-product GEMM)"
timeout 600 python -m pytest tests/test_gpu_parity.py -q --timeout 300 -x -k "cgemm3 or test_gemm_f16x3 or streamk" > gpurun_out/j4_sanity.log 2>&1
tail -5 gpurun_out/j4_sanity.log
grep -q "passed" gpurun_out/j4_sanity.log || { tail -40 gpurun_out/j4_sanity.log; }
echo "== gemm experiments"
( for sk in 1 0 2; do TG_GEMM_STREAMK=$sk timeout 300 python tools/exp_gemm2.py; done
  TG_LIB_PATH=$PWD/temgymcore_b200/libtemgym_b200_ck256.so TG_GEMM_STREAMK=1 timeout 300 python tools/exp_gemm2.py
  TG_LIB_PATH=$PWD/temgymcore_b200/libtemgym_b200_ck256.so TG_GEMM_STREAMK=0 timeout 300 python tools/exp_gemm2.py ) > gpurun_out/j4_gemm.log 2>&1
cat gpurun_out/j4_gemm.log
echo "== tensor path tests"
timeout 1500 python -m pytest tests/test_gpu_parity.py -q --timeout 900 -k "tensor or c3_full or container or host_pipeline" > gpurun_out/j4_pytest.log 2>&1
tail -12 gpurun_out/j4_pytest.log
echo "== bench headline both formulations"
for g in 1 0; do TG_TENSOR_GAUSS=$g timeout 600 python bench.py --steps 10 --warmup 3 --skip-c3 --no-cpu-baseline 2>/dev/null | grep -E '"section": "(headline|e2e)"' | cut -c1-400; done
