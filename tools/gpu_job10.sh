#!/usr/bin/env bash
mkdir -p gpurun_out
echo "== correctness"
timeout 900 python -m pytest tests/test_gpu_parity.py -q --timeout 600 -k "gemm or streamk or cgemm3 or host_pipeline or emulated_ranks or tensor_path_parity or gather or culling" 2>&1 | tail -4
echo "== gemm timings: plain split-K (1) vs head/tail (3)"
for sk in 1 3; do TG_GEMM_STREAMK=$sk timeout 300 python tools/exp_gemm2.py; done 2>&1 | grep -v "^$"
echo "== bench headline, 4-mult vs 3-product"
for g in 0 1; do TG_TENSOR_GAUSS=$g timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | grep -E '"section": "(headline|e2e|c3_biprism)"' | cut -c1-900; done
