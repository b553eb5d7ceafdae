#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sanitizer.py -q --timeout 900 -k "gemm or streamk or cgemm3 or tensor or host_pipeline or emulated_ranks or c2_full or c3_full or decompose or sanitizer or peer_image" > gpurun_out/j12_pytest.log 2>&1
tail -6 gpurun_out/j12_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --skip-c3 --no-cpu-baseline 2>/dev/null | grep -E '"section": "(headline|roofline_tensor_path|e2e)"' | cut -c1-700
