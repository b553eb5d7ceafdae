#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q --timeout 600 -k "derivatives or krivanek or jets" > gpurun_out/j16_pytest.log 2>&1
tail -15 gpurun_out/j16_pytest.log
timeout 300 python tools/exp_jets.py 2>&1 | tee gpurun_out/j16_jets.log
for tool in memcheck racecheck synccheck; do
  timeout 600 /usr/local/cuda/bin/compute-sanitizer --tool $tool python tests/sanitizer_subset.py jets > gpurun_out/j16_san_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "SUMMARY|section" gpurun_out/j16_san_$tool.log
done
