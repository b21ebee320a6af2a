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
    # tree optimisation inside set_tris (k_parents, k_treelet; 3 passes) and wide nodes: same hits
    b.set_tree_optimisation(1, 3)
    b.set_wide_nodes(1)
    b.set_tris(tris)
    h2 = b.nearest_hits(rays)
    for f in h.dtype.names:
        assert np.array_equal(h[f], h2[f], equal_nan=True), f
    # opt-in watertight kernels on the optimised tree
    b.set_triangle_test(1)
    b.set_tris(tris)
    w = b.nearest_hits(rays)
    assert (w["valid"] != h["valid"]).sum() <= 2
    b.nearest_hits(rays[:5000], "valid")
    b.set_triangle_test(0)
    b.set_wide_nodes(2)
    b.set_tree_optimisation(3, 2)
print("sanitizer workload ok")
PY
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san.py > gpurun_out/sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; tail -4 gpurun_out/sanitizer_$tool.log
done
