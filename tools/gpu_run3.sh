#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for fb in 1 0; do
PRT_B200_FAST_BOXES=$fb timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2_fb$fb.json 2> gpurun_out/bench_c2_fb$fb.err; echo "bench fb=$fb rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_c2_fb$fb.json')); print({k:d[k] for k in ('value','ms_per_step','build_mtris_s','gpu_launches','clocks')}, d['e2e']['value'], d['roofline']['frac_of_l2'], d['roofline']['nodes_per_ray'], d['roofline']['tris_per_ray'])"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 5 -c 1 -f -o gpurun_out/prof_trace_c2_v2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_trace.log 2>&1; echo "ncu trace rc=$?"
