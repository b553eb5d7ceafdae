#!/usr/bin/env bash
# tile-binned sum, fifth pass: beamlet-side marking (atomicOr bitmaps) instead of per-tile scans
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q --timeout 900 -k "tensor_binned or c3_full_size or peer_stores_emulated" > gpurun_out/j25_pytest.log 2>&1
tail -5 gpurun_out/j25_pytest.log
timeout 300 python tools/exp_binned.py quick 2>&1 | tail -1 | tee gpurun_out/j25_binned.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bin_|binned|gemm_x3" -c 60 --csv --log-file gpurun_out/j25_launches.csv python tools/exp_binned.py quick > /dev/null 2>&1
python tools/summarize_ncu.py launches gpurun_out/j25_launches.csv gpurun_out/j25_launches.md; sed -n 5,16p gpurun_out/j25_launches.md
