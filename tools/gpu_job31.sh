#!/usr/bin/env bash
# whole GPU tier + smoke on the final build (the ray kernel gained its bounded-grid loop), short bench for the ray numbers
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 2400 python -m pytest tests -x -q -m gpu --timeout 1200 > gpurun_out/final2_pytest.log 2>&1
tail -4 gpurun_out/final2_pytest.log
timeout 900 python bench.py --no-cpu-baseline --skip-c3 > gpurun_out/final2_bench.json 2> gpurun_out/final2_bench.err
grep -E '"section": "rays"' gpurun_out/final2_bench.json | cut -c1-1200; tail -n 1 gpurun_out/final2_bench.json | cut -c1-300
