#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
bash tools/gpu_build_bench.sh 2>&1 | grep "^build"
show() { python -c "
import json,sys; d=json.load(open('$1')); r=d['roofline']; print('$2', {k:round(d[k],4) if isinstance(d[k],float) else d[k] for k in ('value','ms_per_step','build_mtris_s')})"; }
for cfg in c3b; do
timeout 900 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err || tail -5 gpurun_out/bench_$cfg.err; show gpurun_out/bench_$cfg.json $cfg
done
