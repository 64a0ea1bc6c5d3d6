#!/bin/bash
# Build libfbdsp.so for sm_100a in-tree (travels to the GPU box with the snapshot).
set -e
cd "$(dirname "$0")"
mkdir -p lib
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
OUT=${FB_OUT:-lib/libfbdsp.so}
# (into a temporary name first: a snapshot taken while the build runs must never see a half-written library)
$NVCC $FLAGS ${FB_PTXAS_V:+-Xptxas -v} ${NVCC_EXTRA} -shared -cudart static -o $OUT.tmp csrc/*.cu
mv -f $OUT.tmp $OUT
echo "built $(pwd)/${FB_OUT:-lib/libfbdsp.so}"
