#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_c2.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c2.json')); print({k:d[k] for k in ('value','ms_per_step','build_mtris_s','gpu_launches','clocks')}); print(d['e2e']); print(d['cpu_baseline'])"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_c2_ref.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_c2_ref.json')); print('ref arm', d['value'], d['unit'], d['cpu_baseline']['cores'], d['build_mtris_s'])"
