#!/bin/bash
# Build a variant of libmgb200.so with extra nvcc flags into learnmultigrid_b200/_variants/ (kernel experiments:
# tools/gpu_variants.sh swaps them in on the GPU box).   usage: tools/build_variant.sh NAME -DMGB_SELL_BLOCK=128 ...
set -e
name=$1; shift
cd "$(dirname "$0")/../learnmultigrid_b200/csrc"
out=build_$name
mkdir -p $out ../_variants
pids=()
for f in *.cu; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v \
      --expt-relaxed-constexpr "$@" -c $f -o $out/${f%.cu}.o 2> $out/${f%.cu}.ptxas.log &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../_variants/libmgb200_$name.so $out/*.o
echo built ../_variants/libmgb200_$name.so
