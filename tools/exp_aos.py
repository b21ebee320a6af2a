"""SoA vs AoS traversal kernels on C2, whole batch vs pipeline-sized chunks (device-timed)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import portablert_b200 as prt
from portablert_b200 import hitreg, scenes

b = prt.CUDABackend(device=0)
prt.select_backend(b)
tris = scenes.blob()
rays = scenes.pinhole_rays(1920, 1080, cam=(0.0, 0.0, -0.3))
b.set_tris(tris)
d_rays = torch.from_numpy(rays).cuda()
n = len(rays)
uv = torch.empty(n, 2, device="cuda"); t = torch.empty(n, device="cuda")
pid = torch.empty(n, dtype=torch.int32, device="cuda"); p = torch.empty(n, 3, device="cuda")
valid = torch.empty(n, dtype=torch.uint8, device="cuda")
aos = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(fn, reps=7):
    ms = []
    for _ in range(reps):
        flush.zero_()
        torch.cuda.synchronize()
        ms.append(fn())
    return float(np.median(ms[2:]))


for mask in (hitreg.ALL, hitreg.T | hitreg.PID):
    for cnt in (n, 1 << 19, 393216, 1 << 18, 1 << 17, 1 << 16):
        off = (n - cnt) // 2 // 1920 * 1920  # middle of the frame
        ptr = d_rays.data_ptr() + off * 24
        soa = run(lambda: b.trace_dev(ptr, cnt, mask, uv.data_ptr(), t.data_ptr(), pid.data_ptr(),
                                      p.data_ptr(), valid.data_ptr()))
        ao = run(lambda: b.trace_dev_aos(ptr, cnt, mask, aos.data_ptr()))
        print("mask %2d rays %8d  SoA %.4f ms (%.0f Mrays/s)   AoS %.4f ms (%.0f Mrays/s)"
              % (mask, cnt, soa, cnt / soa / 1e3, ao, cnt / ao / 1e3))
