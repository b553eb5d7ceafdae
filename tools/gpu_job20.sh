#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python tools/exp_gemm3.py 2>&1 | tee gpurun_out/j20_gemm3.log
