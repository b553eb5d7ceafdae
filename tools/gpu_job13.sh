#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q --timeout 600 -k "plan or peer_image or emulated" 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 --skip-c3 --no-cpu-baseline 2>/dev/null | grep -E '"section": "(headline|e2e)"' | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --skip-c3 --no-cpu-baseline 2>/dev/null | tail -n 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['gpu_launches'], d['config']['method'], d['ms_per_step'])"
