#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sanitizer.py -q --timeout 900 -k "gather or culling or c3_full or field or cost or sanitizer or emulated" > gpurun_out/j14_pytest.log 2>&1
tail -6 gpurun_out/j14_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | grep -E '"section": "(headline|c3_biprism)"' | cut -c1-2300
