#!/bin/bash
mkdir -p gpurun_out
PRT_B200_WIDE=1 timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu_wide.log 2>&1; echo "pytest(wide=1) rc=$?"; tail -3 gpurun_out/pytest_gpu_wide.log
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest(default) rc=$?"; tail -3 gpurun_out/pytest_gpu.log
show() { python -c "
import json,sys; d=json.load(open('$1')); r=d['roofline']; print('$2', 'trace_ms',round(d['ms_per_step'],4),'Mrays/s',round(d['value']),'build_ms',round(d['build']['ms'],4),'nodes',round(r['nodes_per_ray'],2),'tris',round(r['tris_per_ray'],2))"; }
export PRT_BENCH_C4_RAYS=10000000
for cfg in c2 c3 c3b c5 c4; do
for w in 0 1 2; do
PRT_B200_WIDE=$w timeout 900 python bench.py --config $cfg --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/v.json 2> gpurun_out/v.err || tail -3 gpurun_out/v.err; show gpurun_out/v.json "$cfg wide=$w"
done
done
