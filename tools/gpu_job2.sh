#!/usr/bin/env bash
mkdir -p gpurun_out
echo "== clock experiment"
timeout 300 python tools/exp_clock.py > gpurun_out/j2_clock.log 2>&1; cat gpurun_out/j2_clock.log | tail -4
echo "== gemm experiments (default policy)"
timeout 300 python tools/exp_gemm.py > gpurun_out/j2_gemm.log 2>&1; cat gpurun_out/j2_gemm.log
echo "== e2e"
( for br in 256 512; do TG_E2E_BLOCK_ROWS=$br timeout 200 python tools/exp_e2e2.py packed 2>&1 | tail -1; done ) > gpurun_out/j2_e2e.log 2>&1
cat gpurun_out/j2_e2e.log
echo "== pytest"
timeout 2700 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/j2_pytest.log 2>&1
tail -25 gpurun_out/j2_pytest.log
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/j2_bench.json 2> gpurun_out/j2_bench.err
tail -n 1 gpurun_out/j2_bench.json | cut -c1-6000; tail -5 gpurun_out/j2_bench.err
