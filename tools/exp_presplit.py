"""CPU experiment (host-driven kernel source, tests/emu): would pre-splitting large triangles help
the interior scene once the tree is treelet-optimised?  Cost estimate only (sub-triangles become
triangles of their own).  Result (58 764-triangle interior, 2 passes): boxes per primary ray
17.1 -> 24.0 / 24.8 / 25.5 when triangles above 400x / 100x / 25x the median box area are
midpoint-subdivided, triangles per ray 2.22 -> 2.17: no -- the optimiser already lifts large
triangles to the top of the tree, and splitting them only adds leaves in empty space."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from emu import Emu  # noqa: E402
from portablert_b200 import scenes  # noqa: E402


def split(tris, thresh):
    out, work = [], tris.reshape(-1, 3, 3)
    while len(work):
        d = work.max(1) - work.min(1)
        a = d[:, 0] * d[:, 1] + d[:, 1] * d[:, 2] + d[:, 2] * d[:, 0]
        big = a > thresh
        out.append(work[~big])
        w = work[big]
        if not len(w):
            break
        m01, m12, m20 = (w[:, 0] + w[:, 1]) * 0.5, (w[:, 1] + w[:, 2]) * 0.5, (w[:, 2] + w[:, 0]) * 0.5
        work = np.concatenate([np.stack([w[:, 0], m01, m20], 1), np.stack([m01, w[:, 1], m12], 1),
                               np.stack([m20, m12, w[:, 2]], 1), np.stack([m01, m12, m20], 1)])
    return np.concatenate(out).reshape(-1, 9).astype(np.float32)


if __name__ == "__main__":
    tris = scenes.interior(60000)
    rays = scenes.camera_rays(240, 135, (2, 6, 3), (28, 4, 15))
    lo, hi = tris.reshape(-1, 3).min(0), tris.reshape(-1, 3).max(0)
    inc = scenes.incoherent_rays(20000, lo, hi, 3)
    d = tris.reshape(-1, 3, 3).max(1) - tris.reshape(-1, 3, 3).min(1)
    med = np.median(d[:, 0] * d[:, 1] + d[:, 1] * d[:, 2] + d[:, 2] * d[:, 0])
    for k in (np.inf, 400, 100, 25):
        t2 = split(tris, med * k)
        e = Emu().build(t2, 13)
        e.treelet(2)
        print("split above %sx median area:" % k, len(t2), "tris; boxes, tris per primary ray",
              e.trace(rays)["counts"].mean(0).round(2), "per incoherent ray",
              e.trace(inc)["counts"].mean(0).round(2))
