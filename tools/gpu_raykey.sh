#!/bin/bash
# ray sort key resolution (origin bits, direction bits per axis) on the incoherent configs
mkdir -p gpurun_out
for cfg in ${CFGS:-c3b c4}; do
for k in ${KEYS:-4,4 5,3 6,2 5,5 6,4 7,3 8,2}; do
  PRT_B200_RAYKEY=$k PRT_BENCH_C4_RAYS=${C4_RAYS:-30000000} timeout 900 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/rk.json 2> gpurun_out/rk.err || tail -3 gpurun_out/rk.err
  python -c "
import json; d=json.load(open('gpurun_out/rk.json')); print('$cfg key=$k', 'Mrays/s', round(d['value']), 'ms', round(d['ms_per_step'],3))"
done
done
