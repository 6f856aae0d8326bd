#!/usr/bin/env bash
# Builds libmmdk.so in-tree for sm_100a.  Usage: build.sh [extra nvcc flags]
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC $ARCH --expt-relaxed-constexpr"
mkdir -p ../_build
SRCS="api unet unet_tc unet_fused guide post chain"
pids=()
for f in $SRCS; do
  rm -f ../_build/$f.o            # a failed compile must never leave a stale object for the link step
  extra=""
  # guide.cu / post.cu reproduce the reference's op-by-op fp32 arithmetic: no FMA contraction
  if [ "$f" = guide ] || [ "$f" = post ]; then extra="-fmad=false"; fi
  $NVCC $COMMON $extra -c $f.cu -o ../_build/$f.o "$@" &
  pids+=($!)
done
for pid in "${pids[@]}"; do wait "$pid"; done   # `wait <pid>` propagates each compile's exit status to set -e
OBJS=""
for f in $SRCS; do OBJS="$OBJS ../_build/$f.o"; done
$NVCC -shared $ARCH -o ../libmmdk.so $OBJS -lcudart
echo "built $(cd .. && pwd)/libmmdk.so"
