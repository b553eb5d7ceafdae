#!/usr/bin/env bash
# Builds libtemgym_b200_xla.so (the jax.ffi handlers) when jax is importable; a no-op otherwise.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
root="$(cd "$here/../../.." && pwd)"
inc="$(python -c 'import jax.ffi; print(jax.ffi.include_dir())' 2>/dev/null || true)"
if [ -z "$inc" ]; then
  echo "jax is not importable: the XLA FFI shim is not built (the C ABI + Python surface is the boundary)"
  exit 0
fi
g++ -O2 -std=c++17 -fPIC -shared -I"$inc" -I"$root/include" -I/usr/local/cuda/include \
    "$here/xla_ffi_shim.cc" -L"$root/temgymcore_b200" -ltemgym_b200 -Wl,-rpath,'$ORIGIN' \
    -o "$root/temgymcore_b200/libtemgym_b200_xla.so"
echo "built $root/temgymcore_b200/libtemgym_b200_xla.so"
