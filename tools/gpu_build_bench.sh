#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python - <<'PY'
import numpy as np, torch, sys
sys.path.insert(0,'.')
import portablert_b200 as prt
from portablert_b200 import scenes
prt.select_backend(prt.cuda_backend)
b=prt.cuda_backend
for name,tris in (("69k blob",scenes.blob()),("262k interior",scenes.interior()),("1M heightfield",scenes.heightfield(0)),("10M spheres",scenes.sphere_field(10000))):
    d=torch.from_numpy(tris).cuda(); torch.cuda.synchronize()
    ms=[b.set_tris_dev(d.data_ptr(),len(tris)) for _ in range(8)]
    m=float(np.mean(ms[3:]))
    print(f"build {name:16s} {len(tris):9d} tris  {m:.3f} ms  {len(tris)/m/1e3:.0f} Mtris/s")
PY
