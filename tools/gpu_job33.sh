#!/usr/bin/env bash
# tiled operands on the dense 3-product path: parity subset + A/B timing of the C2 step
mkdir -p gpurun_out
for t in 1 0 1 0; do TG_GEMM_TILED=$t timeout 200 python tools/exp_c2.py 2>&1 | tail -1; done | tee gpurun_out/j33_c2.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q --timeout 600 -k "tensor_path or cgemm3 or c2_full_size_tensor or host_pipeline_row or gaussian_image_plan or peer_stores or peer_image_world or fp16_dynamic" > gpurun_out/j33_pytest.log 2>&1
tail -3 gpurun_out/j33_pytest.log
