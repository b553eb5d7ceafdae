#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q --timeout 600 -k "emulated_ranks or peer_image_world_size_one or gather or culling" > gpurun_out/j9_pytest.log 2>&1
tail -15 gpurun_out/j9_pytest.log
