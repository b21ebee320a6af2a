#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/b.py <<'PY'
import numpy as np, torch, sys
sys.path.insert(0,'.')
import portablert_b200 as prt
from portablert_b200 import scenes
prt.select_backend(prt.cuda_backend)
b=prt.cuda_backend
n=int(sys.argv[1])
tris=scenes.sphere_field(n)
d=torch.from_numpy(tris).cuda(); torch.cuda.synchronize()
for _ in range(3): b.set_tris_dev(d.data_ptr(),len(tris))
PY
for n in 1000 10000; do
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"^k_" --csv --log-file gpurun_out/build_launches_$n.csv python /tmp/b.py $n > /dev/null 2>&1; echo "rc=$?"
done
