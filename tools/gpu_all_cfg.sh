#!/bin/bash
# GPU tests + device-timed numbers for all configs (C2, C3, C3 bounce, C5, C4 full)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
show() { python -c "
import json,sys; d=json.load(open('$1')); r=d['roofline']; print('$2', {k:round(d[k],4) if isinstance(d[k],float) else d[k] for k in ('value','ms_per_step','build_mtris_s')}, 'build_ms',round(d['build']['ms'],3),'e2e',round(d['e2e']['value'],1), 'fracL2',round(r['frac_of_l2'],3), 'nodes',round(r['nodes_per_ray'],2), 'tris',round(r['tris_per_ray'],2))"; }
for cfg in c2 c3 c3b c5; do
timeout 900 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err || tail -5 gpurun_out/bench_$cfg.err; show gpurun_out/bench_$cfg.json $cfg
done
if [ "$1" = "c4" ]; then
timeout 1500 python bench.py --config c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err || tail -5 gpurun_out/bench_c4.err; show gpurun_out/bench_c4.json c4
fi
