#!/bin/bash
# treelet SAH optimisation: GPU test, then build ms / Mrays/s / boxes per ray with 0..3 passes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -s -k "treelet or watertight" > gpurun_out/pytest_treelet.log 2>&1; echo "pytest rc=$?"; grep -E "pass\(es\)|passed|failed|Error|leaks" gpurun_out/pytest_treelet.log | tail -20
for cfg in ${CFGS:-c2 c3 c3b c5}; do
  for p in ${PASSES:-0 1 2 3}; do
    PRT_B200_TREELET_MODE=$([ $p = 0 ] && echo 0 || echo 1) PRT_B200_TREELET_PASSES=$([ $p = 0 ] && echo 1 || echo $p) timeout 600 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/tl_${cfg}_$p.json 2> gpurun_out/tl.err || tail -3 gpurun_out/tl.err
    python -c "
import json; d=json.load(open('gpurun_out/tl_${cfg}_$p.json')); r=d['roofline']; print('$cfg passes=$p', 'Mrays/s', round(d['value']), 'ms', round(d['ms_per_step'],4), 'build ms', round(d['build']['ms'],3), 'boxes/ray', round(r['nodes_per_ray'],2), 'tris/ray', round(r['tris_per_ray'],2), 'e2e', round(d['e2e']['value']))"
  done
done
