#!/usr/bin/env bash
# tile-binned tensor-core sum: parity tests, C3 timing, per-kernel launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q --timeout 900 -k "tensor_binned or c3_full_size or auto_dispatch or streamk_is_deterministic or tensor_path_parity" > gpurun_out/j21_pytest.log 2>&1
tail -25 gpurun_out/j21_pytest.log
timeout 600 python tools/exp_binned.py 2>&1 | tee gpurun_out/j21_binned.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/j21_launches.csv python tools/exp_binned.py > /dev/null 2>&1
python tools/summarize_ncu.py launches gpurun_out/j21_launches.csv gpurun_out/j21_launches.md; head -30 gpurun_out/j21_launches.md
