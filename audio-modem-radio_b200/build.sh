#!/bin/bash
# Build libfbdsp.so for sm_100a in-tree (travels to the GPU box with the snapshot).
set -e
cd "$(dirname "$0")"
mkdir -p lib
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2 --use_fast_math=false"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
$NVCC $FLAGS ${FB_PTXAS_V:+-Xptxas -v} ${NVCC_EXTRA} -shared -cudart static -o ${FB_OUT:-lib/libfbdsp.so} csrc/*.cu -lcufft
echo "built $(pwd)/${FB_OUT:-lib/libfbdsp.so}"
