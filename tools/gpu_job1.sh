#!/usr/bin/env bash
# first GPU job of round 2: sanity of the new GEMM, then tests, experiments, sanitizer subset, short bench
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/j1_smi.txt 2>&1
echo "== sanity" > gpurun_out/j1_sanity.log
timeout 300 python tests/sanitizer_subset.py gemm field >> gpurun_out/j1_sanity.log 2>&1
echo "rc=$?" >> gpurun_out/j1_sanity.log
if ! grep -q "section field: ok" gpurun_out/j1_sanity.log; then echo "SANITY FAILED"; tail -30 gpurun_out/j1_sanity.log; exit 1; fi
echo "== gemm experiments"
TG_GEMM_STREAMK=0 timeout 300 python tools/exp_gemm.py > gpurun_out/j1_gemm.log 2>&1
TG_GEMM_STREAMK=1 timeout 300 python tools/exp_gemm.py >> gpurun_out/j1_gemm.log 2>&1
cat gpurun_out/j1_gemm.log
echo "== e2e experiments"
( TG_E2E_BLOCK_ROWS=-1 TG_HOST_TIMING=1 timeout 200 python tools/exp_e2e2.py pcie 2>&1 | tail -8
  for br in 128 256 512; do TG_E2E_BLOCK_ROWS=$br timeout 200 python tools/exp_e2e2.py packed 2>&1 | tail -1; done
  TG_E2E_BLOCK_ROWS=256 timeout 200 python tools/exp_e2e2.py 2>&1 | tail -1
  TG_E2E_BLOCK_ROWS=256 TG_HOST_TIMING=1 timeout 200 python tools/exp_e2e2.py packed 2>&1 | tail -4 ) > gpurun_out/j1_e2e.log 2>&1
cat gpurun_out/j1_e2e.log
echo "== pytest"
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/j1_pytest.log 2>&1
tail -25 gpurun_out/j1_pytest.log
echo "== sanitizer"
for tool in racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tests/sanitizer_subset.py all > gpurun_out/j1_san_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "SUMMARY|section|hazard|Error" gpurun_out/j1_san_$tool.log | head -20
done
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/j1_bench.json 2> gpurun_out/j1_bench.err
tail -c 3000 gpurun_out/j1_bench.json; tail -5 gpurun_out/j1_bench.err
