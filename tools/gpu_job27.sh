#!/usr/bin/env bash
# tile-binned sum: host-pipeline test; factor kernel occupancy / prefetch variants (experiment builds via TG_LIB_PATH)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q --timeout 900 -k "tensor_binned or peer_stores_emulated" > gpurun_out/j27_pytest.log 2>&1
tail -3 gpurun_out/j27_pytest.log
for lib in libtemgym_b200 libtg_exp_mb10 libtg_exp_pf libtg_exp_mb10pf; do
  echo "== $lib"; TG_LIB_PATH=$PWD/temgymcore_b200/$lib.so timeout 300 python tools/exp_binned.py quick 2>&1 | tail -1
done | tee gpurun_out/j27_variants.log
