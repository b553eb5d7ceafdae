#!/usr/bin/env bash
# round-2 profiles: launch list of the bench command, full captures of the culled field kernel on C3, the stream-K
# GEMM on a row shard, and the CTA-pair kernel
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/r2_launches_bench.csv)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"field_grid_kernel" -s 2 -c 1 -o gpurun_out/r2_field_c3 python tools/prof_kernels.py field_c3 > gpurun_out/r2_p_field.log 2>&1; echo "field rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_x3_kernel" -s 2 -c 1 -o gpurun_out/r2_gemm_shard128 python tools/prof_kernels.py gemm_shard > gpurun_out/r2_p_gs.log 2>&1; echo "gemm shard rc=$?"
TG_GEMM_PAIR=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_x3_pair" -s 2 -c 1 -o gpurun_out/r2_gemm_pair python tools/prof_kernels.py gemm > gpurun_out/r2_p_pair.log 2>&1; echo "pair rc=$?"
ls -la gpurun_out/*.ncu-rep
