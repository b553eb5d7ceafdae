#!/usr/bin/env bash
# ray kernel with a bounded grid (TG_TRACE_PERSIST CTAs per SM): C1 at 1e6 / 1e7 rays, C4 at 1e7
mkdir -p gpurun_out
for p in 0 6 12 4 8; do TG_TRACE_PERSIST=$p timeout 300 python tools/exp_rays.py 2>&1 | grep PERSIST; done | tee gpurun_out/j30_rays.log
