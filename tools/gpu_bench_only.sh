#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 "$@" > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_c2.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c2.json')); print({k:d[k] for k in ('value','ms_per_step','build_mtris_s','gpu_launches','clocks')}); print(d['e2e']); print(d['cpu_baseline']); print({k:d['roofline'][k] for k in ('achieved','frac','frac_of_l2','bytes_per_ray','l2_read_gbs_measured','hbm_read_gbs_measured')})"
