#!/usr/bin/env bash
# Builds libmmdk.so in-tree for sm_100a.  Usage: build.sh [extra nvcc flags]
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC $ARCH --expt-relaxed-constexpr"
mkdir -p ../_build
$NVCC $COMMON -c api.cu -o ../_build/api.o "$@" &
$NVCC $COMMON -c unet.cu -o ../_build/unet.o "$@" &
$NVCC $COMMON -c unet_tc.cu -o ../_build/unet_tc.o "$@" &
$NVCC $COMMON -fmad=false -c guide.cu -o ../_build/guide.o "$@" &
wait
$NVCC -shared $ARCH -o ../libmmdk.so ../_build/api.o ../_build/unet.o ../_build/unet_tc.o ../_build/guide.o -lcudart
echo "built $(cd .. && pwd)/libmmdk.so"
