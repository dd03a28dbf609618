#!/bin/bash
# Builds libmdk.so in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
out="$here/../libmdk.so"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
     -Xcompiler -fPIC -Xcompiler -O2 -Xcompiler -pthread -shared ${MDK_NVCC_EXTRA:-} \
     -o "$out" "$here/common.cu" "$here/rdf.cu" "$here/dynamics.cu" "$here/transform.cu" "$here/flux.cu" "$here/rdf_sort.cu" "$here/ingest.cpp"
echo "built $out"
