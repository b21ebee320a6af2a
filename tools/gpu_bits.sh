#!/bin/bash
mkdir -p gpurun_out
show() { python -c "
import json,sys; d=json.load(open('$1')); r=d['roofline']; print('$2', 'trace_ms',round(d['ms_per_step'],4),'build_ms',round(d['build']['ms'],4),'nodes',round(r['nodes_per_ray'],2),'tris',round(r['tris_per_ray'],2))"; }
export PRT_BENCH_C4_RAYS=10000000
for bits in 10 13 16 21; do
PRT_B200_MORTON_BITS=$bits timeout 900 python bench.py --config c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/v.json 2> gpurun_out/v.err || tail -3 gpurun_out/v.err; show gpurun_out/v.json "c4(10M tris,10M rays) bits=$bits"
done
