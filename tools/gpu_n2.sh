#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L > gpurun_out/gpus.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_c2_n$N.json 2> gpurun_out/bench_c2_n$N.err
echo "rc=$?" | tee gpurun_out/bench_c2_n$N.rc
cat gpurun_out/gpus.txt
wc -c gpurun_out/bench_c2_n$N.json
grep -v "^W\|Warning\|warnings" gpurun_out/bench_c2_n$N.err | tail -20
python -c "
import json; d=json.load(open('gpurun_out/bench_c2_n$N.json')); print({k:d[k] for k in ('value','n_gpus','ms_per_step','build_mtris_s','gpu_launches','clocks')}, d['e2e'])"
