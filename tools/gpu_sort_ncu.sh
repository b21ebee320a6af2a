#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/b.py <<'PY'
import numpy as np, torch, sys
sys.path.insert(0,'.')
import portablert_b200 as prt
from portablert_b200 import scenes
prt.select_backend(prt.cuda_backend)
b=prt.cuda_backend
tris=scenes.sphere_field(10000)
d=torch.from_numpy(tris).cuda(); torch.cuda.synchronize()
for _ in range(2): b.set_tris_dev(d.data_ptr(),len(tris))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_hierarchy" -s 1 -c 1 -f -o gpurun_out/prof_build_10m python /tmp/b.py > gpurun_out/ncu_build10m.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_build10m.log
