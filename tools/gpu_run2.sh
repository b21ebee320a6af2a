#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_c2.json')); print({k:d[k] for k in ('value','ms_per_step','build_mtris_s','gpu_launches','clocks')}, d['e2e'], d['roofline']['frac_of_l2'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 4 -c 1 -f -o gpurun_out/prof_trace_c2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_trace.log 2>&1; echo "ncu trace rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_sort_scatter|k_refit|k_leaves|k_karras" -s 24 -c 9 -f -o gpurun_out/prof_build_c2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_build.log 2>&1; echo "ncu build rc=$?"
ls -la gpurun_out
