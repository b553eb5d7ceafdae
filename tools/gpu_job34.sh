#!/usr/bin/env bash
# confirmation of the default build (row-major dense operands): smoke + a quick tensor-path subset
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q --timeout 150 -k "tensor_path_parity or gaussian_image_plan_cuda or tensor_path_rows" 2>&1 | tail -2
