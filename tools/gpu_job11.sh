#!/usr/bin/env bash
mkdir -p gpurun_out
echo "== full GPU test tier"
timeout 2700 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/j11_pytest.log 2>&1
tail -8 gpurun_out/j11_pytest.log
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/j11_bench.json 2> gpurun_out/j11_bench.err
grep -E '"section": "(headline|roofline_tensor_path|e2e|c3_biprism|cpu_baseline)"' gpurun_out/j11_bench.json | cut -c1-1500; tail -3 gpurun_out/j11_bench.err
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-400
echo "== chunk 256 experiment (plain split-K)"
TG_LIB_PATH=$PWD/temgymcore_b200/libtemgym_b200_ck256.so timeout 300 python tools/exp_gemm2.py 2>&1 | grep -v "^$"
timeout 300 python tools/exp_gemm2.py 2>&1 | grep -v "^$"
