#!/bin/bash
# compute-sanitizer passes on small inputs (memcheck: out-of-bounds / misaligned; racecheck: shared
# memory hazards in the sort and hierarchy kernels; synccheck)
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import numpy as np, sys
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import portablert_b200 as prt
from portablert_b200 import scenes
from oracle import Oracle
import parity
prt.select_backend(prt.cuda_backend)
b = prt.cuda_backend
orc = Oracle()
for tris in (scenes.KAT_TRI, scenes.blob(24, 24), scenes.interior(4000), np.repeat(scenes.blob(6, 6), 20, axis=0)):
    lo, hi = tris.reshape(-1,3).min(0), tris.reshape(-1,3).max(0)
    rays = np.concatenate([scenes.pinhole_rays(64, 48), scenes.incoherent_rays(70000, lo - 0.1, hi + 0.1, 3)])
    b.set_tris(tris)
    b.set_ray_sorting(1)
    h = b.nearest_hits(rays)
    orc.build(tris)
    rep = parity.compare(orc.trace(rays), parity.from_structured(h), tris, rays, orc, max_replay=10**6)
    parity.assert_parity(rep)
    for tags in (("valid",), ("t", "primitive_id"), ("uv", "p")):
        b.nearest_hits(rays[:5000], *tags)
print("sanitizer workload ok")
PY
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san.py > gpurun_out/sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; tail -4 gpurun_out/sanitizer_$tool.log
done
