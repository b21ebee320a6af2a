#!/bin/bash
# device-timed numbers for the other BASELINE configs (C3 primary, C3 bounce, C5, C4 full size)
mkdir -p gpurun_out
show() { python -c "
import json,sys; d=json.load(open('$1')); r=d['roofline']; print('$2', {k:round(d[k],4) if isinstance(d[k],float) else d[k] for k in ('value','ms_per_step','build_mtris_s')}, 'build_ms',round(d['build']['ms'],3),'e2e',round(d['e2e']['value'],1), 'fracHBM',round(r['frac'],3),'fracL2',round(r['frac_of_l2'],3), 'nodes',round(r['nodes_per_ray'],2), 'tris',round(r['tris_per_ray'],2), 'B/ray', round(r['bytes_per_ray']), d['config']['rays_per_gpu'], d['config']['tris'])"; }
for cfg in c3 c3b c5; do
timeout 900 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err || tail -5 gpurun_out/bench_$cfg.err; show gpurun_out/bench_$cfg.json $cfg
done
timeout 1500 python bench.py --config c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err || tail -5 gpurun_out/bench_c4.err; show gpurun_out/bench_c4.json c4
