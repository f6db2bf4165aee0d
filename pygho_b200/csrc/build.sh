#!/usr/bin/env bash
# Build libpygho_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libpygho_b200.so
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --use_fast_math=false -Xptxas -v"
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC ${EXTRA_FLAGS:-}"
mkdir -p build
objs=()
for f in seg_gmr plan masked mamamm_tc mamamm_smem fused_mlp linear_stats hodata; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ common.cuh -nt build/$f.o ] || [ ticket.cuh -nt build/$f.o ] || [ ../../include/pygho_b200.h -nt build/$f.o ]; then
    $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c $f.cu -o build/$f.o &
  fi
  objs+=(build/$f.o)
done
wait
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o $OUT "${objs[@]}"
echo "built $(realpath $OUT)"
