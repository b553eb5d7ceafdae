#!/usr/bin/env bash
# Builds libtemgym_b200.so (sm_100a only) next to the Python package.
# trace.cu / coeffs.cu: -fmad=false (bit-faithful fp64, see the file headers); field.cu: FMA on.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
root="$(cd "$here/../.." && pwd)"
out="${TG_BUILD_OUT:-$here/../libtemgym_b200.so}"   # TG_BUILD_OUT / TG_BUILD_DEFS: experiment builds
obj="$here/_obj${TG_BUILD_TAG:-}"
mkdir -p "$obj"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I$root/include -I$here $ARCH ${TG_BUILD_DEFS:-}"
VERBOSE="${TG_PTXAS_V:+-Xptxas -v}"
pids=()
$NVCC $COMMON $VERBOSE -fmad=false -c "$here/trace.cu"    -o "$obj/trace.o" & pids+=($!)
$NVCC $COMMON $VERBOSE -fmad=false -c "$here/coeffs.cu"   -o "$obj/coeffs.o" & pids+=($!)
$NVCC $COMMON $VERBOSE -fmad=false -c "$here/stem4d.cu"   -o "$obj/stem4d.o" & pids+=($!)
$NVCC $COMMON $VERBOSE -fmad=false -c "$here/jets.cu"     -o "$obj/jets.o" & pids+=($!)
$NVCC $COMMON $VERBOSE             -c "$here/field.cu"    -o "$obj/field.o" & pids+=($!)
$NVCC $COMMON $VERBOSE             -c "$here/host_api.cu" -o "$obj/host_api.o" & pids+=($!)
$NVCC $COMMON $VERBOSE             -c "$here/separable.cu" -o "$obj/separable.o" & pids+=($!)
$NVCC $COMMON $VERBOSE             -c "$here/peer.cu"     -o "$obj/peer.o" & pids+=($!)
$NVCC $COMMON $VERBOSE -fmad=false -c "$here/sources.cu"  -o "$obj/sources.o" & pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
$NVCC $ARCH -shared -o "$out" "$obj/trace.o" "$obj/coeffs.o" "$obj/field.o" "$obj/host_api.o" "$obj/separable.o" "$obj/stem4d.o" "$obj/jets.o" "$obj/peer.o" "$obj/sources.o" -cudart static
echo "built $out"
