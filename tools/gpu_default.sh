#!/bin/bash
# default settings (lazy tree optimisation): full GPU test suite, smoke, bench lines for every config
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
for cfg in ${CFGS:-c2 c3 c3b c5}; do
  timeout 900 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err || tail -3 gpurun_out/bench_$cfg.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_$cfg.json')); r=d['roofline']; b=d['build']; print('$cfg', 'Mrays/s', round(d['value']), 'ms', round(d['ms_per_step'],4), 'build ms', round(b['ms'],3), 'build+opt ms', round(b['ms_with_optimisation'],3), 'steady set_tris ms', b['set_tris_ms_steady'], 'refits', b['refits_in_timed_steps'], 'height', b['tree_height'], 'boxes/ray', round(r['nodes_per_ray'],2), 'tris/ray', round(r['tris_per_ray'],2), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], 'wt', d.get('watertight'))"
done
