#!/bin/bash
# A/B of a variant build (PRT_B200_LIB=$1) against the in-tree library on every bench config
mkdir -p gpurun_out
V=$1
for cfg in c2 c3 c3b c5; do
  for lib in "" "$V"; do
    PRT_B200_LIB=$lib timeout 600 python bench.py --config $cfg --steps 10 --warmup 3 > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
    python -c "
import json; d=json.load(open('gpurun_out/ab.json')); print('$cfg', '${lib:-default}', 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']))"
  done
done
