"""Host entry point on C2 with pinned buffers: copy-engine pipeline vs zero-copy variants, chunk
sizes; PRT_B200_PIPE_TRACE=1 prints the stage timeline of the last call of each configuration."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import portablert_b200 as prt
from portablert_b200 import hitreg, scenes
from portablert_b200.backend import pinned_empty

tris = scenes.blob()
rays = scenes.pinhole_rays(1920, 1080, cam=(0.0, 0.0, -0.3))
p_rays = pinned_empty(rays.shape, np.float32)
p_rays[...] = rays
ref = None
for zc, ch in ((0, 0), (0, 17), (0, 18), (0, 19)):
    os.environ.pop("PRT_B200_CHUNK_LOG2", None)
    if ch:
        os.environ["PRT_B200_CHUNK_LOG2"] = str(ch)
    os.environ["PRT_B200_PIPE_TRACE"] = "0"
    b = prt.CUDABackend(device=0)
    b.init()
    b.set_tris(tris)
    p_hits = pinned_empty((len(rays),), hitreg.dtype(hitreg.ALL))
    ts = []
    for i in range(40):
        t0 = time.perf_counter()
        b.nearest_hits(p_rays, hitreg.ALL, out=p_hits)
        ts.append((time.perf_counter() - t0) * 1e3)
    got = {f: np.array(p_hits[f]) for f in p_hits.dtype.names}
    if ref is None:
        ref = got
    same = all(np.array_equal(ref[f], got[f], equal_nan=True) for f in ref)
    print("chunk_log2=%2d  wall median %.3f ms  min %.3f ms  (%.0f Mrays/s)  identical=%s"
          % (ch, np.median(ts[5:]), min(ts), len(rays) / np.median(ts[5:]) / 1e3, same), flush=True)
    b.shutdown() if hasattr(b, "shutdown") else None
