#!/usr/bin/env bash
# tile-binned sum: AUTO probe (one read-back); eager timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q --timeout 900 -k "tensor_binned or c3_full_size or auto_dispatch or host_pipeline" > gpurun_out/j28_pytest.log 2>&1
tail -12 gpurun_out/j28_pytest.log
timeout 600 python tools/exp_binned.py 2>&1 | tee gpurun_out/j28_binned.log
