#!/usr/bin/env bash
# memcheck of the tile-binned section (+ gemm), and the sanitizer tier with the new section asserted
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck python tests/sanitizer_subset.py binned gemm > gpurun_out/j32_memcheck.log 2>&1
grep -E "section|ERROR SUMMARY" gpurun_out/j32_memcheck.log
timeout 900 python -m pytest tests/test_gpu_sanitizer.py -x -q --timeout 900 2>&1 | tail -2
