#!/bin/bash
# Builds A/B variants of the library under variants/ (git-ignored; travels to the GPU box):
#   tools/build_variants.sh NAME "EXTRA NVCC FLAGS" [NAME2 "FLAGS2" ...]
# Select one at run time with PRT_B200_LIB=variants/libprt_b200_NAME.so
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/portablert_b200/csrc
mkdir -p $ROOT/variants
while [ $# -ge 2 ]; do
  NAME=$1; FLAGS=$2; shift 2
  B=/tmp/prt_variant_$NAME; mkdir -p $B
  for f in api build sort trace trace_wide trace_exact trace_wt trace_coop; do
    ( nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC \
        --expt-relaxed-constexpr -cudart static $FLAGS -c $SRC/$f.cu -o $B/$f.o ) &
  done
  wait
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o $ROOT/variants/libprt_b200_$NAME.so $B/*.o -ldl
  echo built variants/libprt_b200_$NAME.so
done
