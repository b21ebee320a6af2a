#!/bin/bash
# ncu --set full of the traversal kernel on a given config: tools/gpu_ncu_cfg.sh <cfg> <outname>
mkdir -p gpurun_out
cfg=$1; name=$2
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 3 -c 1 -f -o gpurun_out/$name python bench.py --config $cfg --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$name.log 2>&1; echo "ncu $cfg rc=$?"; tail -2 gpurun_out/ncu_$name.log
