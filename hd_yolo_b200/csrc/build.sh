#!/usr/bin/env bash
# Build the C-ABI shared library for sm_100a, in-tree (hd_yolo_b200/libhdyolo_b200.so).
# -fmad=false: the reference evaluates every product and sum separately in fp32; explicit fmaf
# is used where contraction is wanted (mask contraction).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
root="$(cd "$here/../.." && pwd)"
out="$root/hd_yolo_b200/libhdyolo_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
  -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -shared -cudart static \
  ${HDY_NVCC_EXTRA:-} -I"$root/include" -o "$out" "$here"/*.cu
echo "built $out"
