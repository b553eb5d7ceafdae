#!/usr/bin/env bash
# split-K with tails (mode 1) against plain split-K (mode 4): GEMM timings, then the tensor-path tests and a short bench
mkdir -p gpurun_out
for m in 4 1; do TG_GEMM_STREAMK=$m timeout 300 python tools/exp_gemm2.py 2>&1 | grep -E "M= *(1024|512|256)" ; done | tee gpurun_out/j15_gemm.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -q --timeout 900 -k "gemm or streamk or cgemm3 or tensor or host_pipeline or emulated_ranks or c2_full or plan" > gpurun_out/j15_pytest.log 2>&1
tail -4 gpurun_out/j15_pytest.log
for m in 4 1; do echo "STREAMK=$m"; TG_GEMM_STREAMK=$m timeout 600 python bench.py --steps 20 --warmup 5 --skip-c3 --no-cpu-baseline 2>/dev/null | grep -E '"section": "(headline|roofline_tensor_path|e2e)"' | cut -c1-600; done | tee gpurun_out/j15_bench.log
