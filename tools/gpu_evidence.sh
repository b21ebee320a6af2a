#!/bin/bash
# End-of-round evidence: GPU tests, bench lines for every config, ncu launch lists and full captures.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 --per-mask > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_c2_ref.json 2>/dev/null; echo "ref arm rc=$?"
for cfg in c3 c3b c5; do
timeout 900 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err; echo "bench $cfg rc=$?"
done
timeout 1500 python bench.py --config c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 3 -c 1 -f -o gpurun_out/prof_trace_c2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_trace.log 2>&1; echo "ncu trace c2 rc=$?"
PRT_BENCH_C4_RAYS=10000000 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 3 -c 1 -f -o gpurun_out/prof_trace_c4 python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_trace_c4.log 2>&1; echo "ncu trace c4 rc=$?"
bash tools/gpu_build_ncu.sh
