#!/bin/bash
# Builds libmdk.so in-tree for sm_100a (cross-compiles without a GPU).  One object per
# source, compiled in parallel and rebuilt only when the source (or a header) is newer.
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
out="$here/../libmdk.so"
obj="$here/build"
mkdir -p "$obj"
NVCC_FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17
            -Xcompiler -fPIC -Xcompiler -O2 -Xcompiler -pthread ${MDK_NVCC_EXTRA:-})
srcs=(common.cu rdf.cu dynamics.cu transform.cu flux.cu rdf_sort.cu adf.cu ingest.cpp)
pids=()
objs=()
for s in "${srcs[@]}"; do
  [ -f "$here/$s" ] || continue
  o="$obj/${s%.*}.o"
  objs+=("$o")
  if [ ! -f "$o" ] || [ "$here/$s" -nt "$o" ] || [ "$here/mdk_common.cuh" -nt "$o" ] \
     || [ "$here/../../include/mdk.h" -nt "$o" ] || [ "$here/build.sh" -nt "$o" ]; then
    nvcc "${NVCC_FLAGS[@]}" -c -o "$o" "$here/$s" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -pthread -o "$out" "${objs[@]}"
echo "built $out"
