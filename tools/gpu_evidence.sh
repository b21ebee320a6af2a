#!/bin/bash
# End-of-round evidence: GPU tests, smoke, bench lines for every config, the reference arm, ncu launch
# list and full captures of the traversal and treelet kernels.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 --per-mask > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_c2_ref.json 2>/dev/null; echo "ref arm rc=$?"
for cfg in c3 c3b c5; do
timeout 900 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err; echo "bench $cfg rc=$?"
done
timeout 1500 python bench.py --config c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo "ncu launches rc=$?"
for cfg in c2 c3; do
# (tree optimisation inside set_tris so that launch #3 is the first timed step on the optimised tree)
PRT_B200_TREELET_MODE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 3 -c 1 -f -o gpurun_out/prof_trace_$cfg python bench.py --config $cfg --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_trace_$cfg.log 2>&1; echo "ncu trace $cfg rc=$?"
done
PRT_B200_TREELET_MODE=1 PRT_BENCH_C4_RAYS=10000000 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 3 -c 1 -f -o gpurun_out/prof_trace_c4 python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_trace_c4.log 2>&1; echo "ncu trace c4 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_treelet -s 2 -c 1 -f -o gpurun_out/prof_treelet_c5 python bench.py --config c5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_treelet_c5.log 2>&1; echo "ncu treelet c5 rc=$?"
