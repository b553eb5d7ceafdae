#!/usr/bin/env bash
# tile-binned sum, sixth pass: one fused factor kernel; ncu full captures of the ragged GEMM and the factor kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q --timeout 900 -k "tensor_binned or c3_full_size or peer_stores_emulated" > gpurun_out/j26_pytest.log 2>&1
tail -3 gpurun_out/j26_pytest.log
timeout 300 python tools/exp_binned.py quick 2>&1 | tail -1 | tee gpurun_out/j26_binned.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bin_|binned|gemm_x3|prep|coeffs|trace" -c 90 --csv --log-file gpurun_out/r2_launches_c3_binned.csv python tools/exp_binned.py quick > /dev/null 2>&1
python tools/summarize_ncu.py launches gpurun_out/r2_launches_c3_binned.csv gpurun_out/r2_launches_c3_binned.md; sed -n 5,18p gpurun_out/r2_launches_c3_binned.md
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_x3_kernel" -s 2 -c 1 -o gpurun_out/r2_gemm_ragged_c3 python tools/exp_binned.py quick > gpurun_out/j26_p1.log 2>&1; echo "gemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"factor_binned" -s 2 -c 1 -o gpurun_out/r2_factor_binned_c3 python tools/exp_binned.py quick > gpurun_out/j26_p2.log 2>&1; echo "factor rc=$?"
ls -la gpurun_out/r2_gemm_ragged_c3.ncu-rep gpurun_out/r2_factor_binned_c3.ncu-rep
