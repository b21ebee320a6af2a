#!/bin/bash
# quick iteration: GPU tests (optional) + C2 bench + optional ncu of the SoA trace kernel
mkdir -p gpurun_out
if [ "$1" = "test" ]; then
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
fi
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_c2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c2.json')); print({k:d[k] for k in ('value','ms_per_step','build_mtris_s','gpu_launches','clocks')}, d['e2e']['value'], d['roofline']['frac_of_l2'], d['roofline']['nodes_per_ray'], d['roofline']['tris_per_ray'])"
if [ "$2" = "ncu" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 3 -c 1 -f -o gpurun_out/prof_trace_c2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_trace.log 2>&1; echo "ncu trace rc=$?"
fi
