"""Host entry point with PAGEABLE buffers (what std::vector callers of the reference API pass)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import portablert_b200 as prt
from portablert_b200 import hitreg, scenes

tris = scenes.blob()
rays = scenes.pinhole_rays(1920, 1080, cam=(0.0, 0.0, -0.3))
n = len(rays)
src = np.empty(n * 24, np.uint8); dst = np.empty(n * 24, np.uint8)
for _ in range(3):
    t0 = time.perf_counter(); dst[:] = src; dt = time.perf_counter() - t0
print("single-thread numpy copy of %d MB: %.2f ms (%.1f GB/s)" % (n * 24 >> 20, dt * 1e3, n * 24 / dt / 1e9))
for ch in (0, 19):
    os.environ.pop("PRT_B200_CHUNK_LOG2", None)
    if ch:
        os.environ["PRT_B200_CHUNK_LOG2"] = str(ch)
    os.environ["PRT_B200_PIPE_TRACE"] = os.environ.get("TRACE", "0")
    b = prt.CUDABackend(device=0)
    b.init()
    b.set_tris(tris)
    out = np.zeros((n,), hitreg.dtype(hitreg.ALL))
    for label, fresh in (("reused result array", False), ("fresh result array", True)):
        ts = []
        for i in range(20):
            t0 = time.perf_counter()
            if fresh:
                b.nearest_hits(rays, hitreg.ALL)
            else:
                b.nearest_hits(rays, hitreg.ALL, out=out)
            ts.append((time.perf_counter() - t0) * 1e3)
        print("chunk_log2=%2d pageable in, %s: median %.3f ms (%.0f Mrays/s)" % (
            ch, label, np.median(ts[3:]), n / np.median(ts[3:]) / 1e3), flush=True)
