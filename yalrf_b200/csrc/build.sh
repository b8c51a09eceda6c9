#!/bin/bash
# Builds libyhb.so (sm_100a only) next to the Python package.  -fmad=false: the device models follow the
# reference's separately-rounded scalar arithmetic; the hot loops use explicit fma().
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="${YHB_OUT:-$HERE/../libyhb.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -std=c++17 -O3 -lineinfo -fmad=false -gencode arch=compute_100a,code=sm_100a \
    -Xcompiler -fPIC,-O2,-Wall -shared -o "$OUT" "$HERE/yhb_api.cu" "$@"
echo "built $OUT"
