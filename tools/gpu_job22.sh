#!/usr/bin/env bash
# tile-binned sum, second pass: elliptical footprints, 8 footprints in flight per lane, empty tiles zeroed by a kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q --timeout 900 -k "tensor_binned or c3_full_size or peer_stores_emulated or peer_image_world" > gpurun_out/j22_pytest.log 2>&1
tail -15 gpurun_out/j22_pytest.log
timeout 600 python tools/exp_binned.py 2>&1 | tee gpurun_out/j22_binned.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bin_|binned|gemm_x3" -c 60 --csv --log-file gpurun_out/j22_launches.csv python tools/exp_binned.py > /dev/null 2>&1
python tools/summarize_ncu.py launches gpurun_out/j22_launches.csv gpurun_out/j22_launches.md; head -20 gpurun_out/j22_launches.md
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/j22_bench.json 2> gpurun_out/j22_bench.err
grep -E '"section": "(headline|c3_biprism)"' gpurun_out/j22_bench.json | cut -c1-3500; tail -3 gpurun_out/j22_bench.err
