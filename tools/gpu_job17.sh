#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q --timeout 600 -k "derivatives or krivanek or jets" > gpurun_out/j17_pytest.log 2>&1
tail -5 gpurun_out/j17_pytest.log
timeout 300 python tools/exp_jets.py 2>&1 | tee gpurun_out/j17_jets.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jets_kernel -s 2 -c 1 -o gpurun_out/r2_jets_c4 -f python tools/prof_kernels.py jets_c4 > gpurun_out/j17_ncu.log 2>&1
ls -la gpurun_out/r2_jets_c4.ncu-rep
