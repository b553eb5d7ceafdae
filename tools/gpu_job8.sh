#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q --timeout 600 -k "gather or c3 or culling or field or cost" > gpurun_out/j8_pytest.log 2>&1
tail -5 gpurun_out/j8_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/j8_bench.json 2> gpurun_out/j8_bench.err
grep -E '"section": "(headline|c3_biprism|e2e)"' gpurun_out/j8_bench.json | cut -c1-2600; tail -3 gpurun_out/j8_bench.err
