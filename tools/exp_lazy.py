"""lazy tree optimisation: one-off cost as the library reports it, first scene vs later scenes"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import portablert_b200 as prt
from portablert_b200 import scenes, hitreg
b = prt.cuda_backend; prt.select_backend(b)
for name, tris, rays in (("c2", scenes.blob(), scenes.pinhole_rays(1920, 1080)),
                         ("c3", scenes.interior(), scenes.camera_rays(3840, 2160, (2, 6, 3), (28, 4, 15)))):
    d_tris = torch.from_numpy(tris).cuda(); d_rays = torch.from_numpy(rays).cuda()
    t = torch.empty(len(rays), device="cuda")
    for rep in range(3):
        b.set_tris_dev(d_tris.data_ptr(), len(tris))
        log = []
        for k in range(8):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            ms = b.trace_dev(d_rays.data_ptr(), len(rays), hitreg.T, t=t.data_ptr())
            torch.cuda.synchronize(); wall = (time.perf_counter() - t0) * 1e3
            log.append((round(ms, 3), round(wall, 3), b.tree_depth))
        print(name, "rep", rep, "optimise_ms", round(b.last_optimise_ms, 3), log)
