#!/bin/bash
# A/B of the NUMA-local placement of the shared host batch (N ranks, e2e leg of bench.py)
mkdir -p gpurun_out
N=${1:-8}
{ nvidia-smi topo -m; lscpu | grep -i 'numa\|socket\|model name\|^CPU(s)'; } > gpurun_out/topo.txt 2>&1
for numa in 1 0; do
  PRT_BENCH_NUMA=$numa python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 2951$numa bench.py --gpus $N --steps 10 --warmup 3 \
    > gpurun_out/bench_c2_n${N}_numa$numa.json 2> gpurun_out/bench_c2_n${N}_numa$numa.err
  echo "numa=$numa rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_c2_n${N}_numa$numa.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['config']['numa'])"
done
cat gpurun_out/topo.txt
