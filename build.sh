#!/bin/bash
# Builds libsubgnn_b200.so in-tree (sm_100a only).  Used by __graft_entry__.build().
set -e
cd "$(dirname "$0")"
OUT=subgnn_b200/libsubgnn_b200.so
SRC=$(ls subgnn_b200/csrc/*.cu)
mkdir -p build
OBJS=""
for f in $SRC; do
  o=build/$(basename ${f%.cu}).o
  OBJS="$OBJS $o"
  if [ ! -f $o ] || [ $f -nt $o ] || [ subgnn_b200/csrc/common.cuh -nt $o ] || [ subgnn_b200/csrc/gemm_tile.cuh -nt $o ] || [ subgnn_b200/csrc/lstm_reg.cuh -nt $o ] || [ include/subgnn_b200.h -nt $o ]; then
    rm -f $o; nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --extended-lambda -Xcompiler -fPIC ${NVCC_EXTRA} -c $f -o $o &
  fi
done
wait
for o in $OBJS; do [ -f $o ] || { echo "compile failed: $o"; exit 1; }; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $OUT $OBJS -lcudart
echo built $OUT
