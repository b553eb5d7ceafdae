#!/usr/bin/env bash
# re-validation after the last edits (nb bound, comments): smoke + tile-binned tests + quick C3 timing
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q --timeout 900 -k "tensor_binned or peer_stores_emulated or auto_dispatch" > gpurun_out/j29_pytest.log 2>&1
tail -3 gpurun_out/j29_pytest.log
timeout 300 python tools/exp_binned.py quick 2>&1 | tail -1
