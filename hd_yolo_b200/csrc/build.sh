#!/usr/bin/env bash
# Build the C-ABI shared library for sm_100a, in-tree (hd_yolo_b200/libhdyolo_b200.so).
# -fmad=false: the reference evaluates every product and sum separately in fp32; explicit fmaf
# is used where contraction is wanted (mask contraction).
# Every .cu is compiled to an object of its own (in parallel, only when it or a header changed), then linked.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
root="$(cd "$here/../.." && pwd)"
out="$root/hd_yolo_b200/libhdyolo_b200.so"
obj="$here/.obj"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -Xcompiler -fvisibility=hidden ${HDY_NVCC_EXTRA:-} -I$root/include"
mkdir -p "$obj"
stamp="$(echo "$FLAGS" | md5sum | cut -c1-8)"
newest_header="$(ls -t "$here"/*.cuh "$root"/include/*.h | head -1)"
pids=()
for src in "$here"/*.cu; do
  o="$obj/$(basename "${src%.cu}").$stamp.o"
  if [ ! -f "$o" ] || [ "$src" -nt "$o" ] || [ "$newest_header" -nt "$o" ]; then
    "$NVCC" $FLAGS -c -o "$o" "$src" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
objs=()
for src in "$here"/*.cu; do objs+=("$obj/$(basename "${src%.cu}").$stamp.o"); done
"$NVCC" -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o "$out" "${objs[@]}"
echo "built $out"
