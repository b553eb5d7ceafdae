#!/usr/bin/env bash
# round-2 final records: smoke, the whole GPU test tier, the bench line, the launch list of the bench command
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 2400 python -m pytest tests -x -q -m gpu --timeout 1200 > gpurun_out/final_pytest.log 2>&1
tail -6 gpurun_out/final_pytest.log
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
tail -n 1 gpurun_out/final_bench.json | cut -c1-1500; tail -3 gpurun_out/final_bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -n 1 | cut -c1-600
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/r2_launches_bench.csv)"
