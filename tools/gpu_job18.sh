#!/usr/bin/env bash
# single-launch row-block streaming of the host pipeline; jets with the precomputed lens coefficients
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q --timeout 600 -k "host_pipeline or derivatives or krivanek or jets or make_gaussian_image_host or e2e" > gpurun_out/j18_pytest.log 2>&1
tail -8 gpurun_out/j18_pytest.log
for cfg in "0 256" "1 256" "1 128" "1 512"; do set -- $cfg
  echo "STREAM=$1"; TG_HOST_TIMING=1 TG_E2E_STREAM=$1 TG_E2E_BLOCK_ROWS=$2 timeout 300 python tools/exp_e2e2.py packed 2>&1 | tail -3
done | tee gpurun_out/j18_e2e.log
timeout 300 python tools/exp_jets.py 2>&1 | grep "order 3" | tee gpurun_out/j18_jets.log
