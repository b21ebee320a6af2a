#!/bin/bash
# e2e pipeline of the host entry point: GPU tests, then bench.py at several chunk sizes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for ch in auto 17 18 19; do
  if [ $ch = auto ]; then unset PRT_B200_CHUNK_LOG2; else export PRT_B200_CHUNK_LOG2=$ch; fi
  python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2_ch$ch.json 2> gpurun_out/bench_c2_ch$ch.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_c2_ch$ch.json')); e=d['e2e']; print('chunk_log2=$ch', 'value', round(d['value']), 'e2e', round(e['value']), 'ms', round(e['ms_per_step'],3), 'pageable', round(e['pageable_value']), e.get('pcie'), e.get('frac_of_pcie_floor'))"
done
