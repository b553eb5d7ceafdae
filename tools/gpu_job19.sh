#!/usr/bin/env bash
# host pipeline: flagged per-block launches chained with programmatic dependent launch
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q --timeout 600 -k "host_pipeline or make_gaussian_image_host or e2e or host_buffer or tensor_path_parity" > gpurun_out/j19_pytest.log 2>&1
tail -8 gpurun_out/j19_pytest.log
for cfg in "0 256" "1 256" "1 128" "1 384" "1 512"; do set -- $cfg
  echo "FLAGGED=$1"; TG_HOST_TIMING=1 TG_E2E_FLAGGED=$1 TG_E2E_BLOCK_ROWS=$2 timeout 300 python tools/exp_e2e2.py packed 2>&1 | tail -2
done | tee gpurun_out/j19_e2e.log
TG_E2E_STREAM=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q --timeout 600 -k "host_pipeline" 2>&1 | tail -2
