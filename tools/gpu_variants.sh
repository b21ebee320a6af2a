#!/bin/bash
mkdir -p gpurun_out
show() { python -c "
import json,sys; d=json.load(open('$1')); r=d['roofline']; print('$2', {k:round(d[k],4) if isinstance(d[k],float) else d[k] for k in ('value','ms_per_step')})"; }
export PRT_BENCH_C4_SPHERES=2000 PRT_BENCH_C4_RAYS=8000000
for v in "" variants/libprt_t128_b10.so variants/libprt_t128_b12.so variants/libprt_t64_b16.so variants/libprt_t256_b4.so variants/libprt_t256_b5.so; do
  for cfg in c2 c4; do
    PRT_B200_LIB=$v timeout 600 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/v.json 2> gpurun_out/v.err || tail -3 gpurun_out/v.err; show gpurun_out/v.json "$cfg ${v:-default}"
  done
done
