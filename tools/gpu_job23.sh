#!/usr/bin/env bash
# tile-binned sum, third pass: L2 prefetch of the operand stream, coarse range test + ballot words, packed split
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q --timeout 900 -k "tensor_binned or c3_full_size or peer_stores_emulated" > gpurun_out/j23_pytest.log 2>&1
tail -15 gpurun_out/j23_pytest.log
for pf in 0 3 6 10 16; do echo "TG_BIN_PREFETCH=$pf"; TG_BIN_PREFETCH=$pf timeout 300 python tools/exp_binned.py quick 2>&1 | tail -1; done | tee gpurun_out/j23_prefetch.log
timeout 600 python tools/exp_binned.py 2>&1 | tee gpurun_out/j23_binned.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bin_|binned|gemm_x3" -c 60 --csv --log-file gpurun_out/j23_launches.csv python tools/exp_binned.py quick > /dev/null 2>&1
python tools/summarize_ncu.py launches gpurun_out/j23_launches.csv gpurun_out/j23_launches.md; head -16 gpurun_out/j23_launches.md
