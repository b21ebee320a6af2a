#!/bin/bash
mkdir -p gpurun_out
nproc
for rep in 1 2; do
for lib in variants/libprt_old.so ""; do
  echo "--- lib=${lib:-new}"
  PRT_B200_LIB=$lib timeout 300 python tools/exp_pageable.py 2>&1 | tail -4
done
done
echo "--- new, 16 copy threads"; PRT_B200_COPY_THREADS=16 timeout 300 python tools/exp_pageable.py 2>&1 | tail -4
