#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
show() { python -c "
import json,sys; d=json.load(open('$1')); print('$2', {k:round(d[k],4) if isinstance(d[k],float) else d[k] for k in ('value','ms_per_step','build_mtris_s')}, 'e2e',round(d['e2e']['value'],1), 'fracL2',round(d['roofline']['frac_of_l2'],3), 'nodes',round(d['roofline']['nodes_per_ray'],2), 'tris',round(d['roofline']['tris_per_ray'],2))"; }
for rf in 0 16 24 28; do
PRT_B200_REFILL=$rf timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2_rf$rf.json 2> gpurun_out/err.log || tail -3 gpurun_out/err.log; show gpurun_out/bench_c2_rf$rf.json "c2 refill=$rf"
done
export PRT_BENCH_C4_SPHERES=1000 PRT_BENCH_C4_RAYS=8000000
for rf in 0 16 24 28; do
PRT_B200_REFILL=$rf timeout 600 python bench.py --config c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4s_rf$rf.json 2> gpurun_out/err.log || tail -3 gpurun_out/err.log; show gpurun_out/bench_c4s_rf$rf.json "c4small refill=$rf"
done
