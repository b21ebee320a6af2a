"""How much of the C2 traversal kernel is tail?  Per-ray work distribution, and the kernel time
with the most expensive rays moved to the front of the batch (an oracle schedule)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import portablert_b200 as prt
from portablert_b200 import hitreg, scenes

b = prt.CUDABackend(device=0)
b.init()
b.set_ray_sorting(0)
tris = scenes.blob()
rays = scenes.pinhole_rays(1920, 1080, cam=(0.0, 0.0, -0.3))
b.set_tris(tris)
n = len(rays)
d_rays = torch.from_numpy(rays).cuda()
cnt = torch.zeros(n, 2, dtype=torch.int32, device="cuda")
b.trace_count_dev(d_rays.data_ptr(), n, cnt.data_ptr())
c = cnt.cpu().numpy().astype(np.int64)
steps = c[:, 0] + c[:, 1]
print("steps/ray: mean %.1f  p50 %d  p90 %d  p99 %d  p99.9 %d  p99.99 %d  max %d" % (
    steps.mean(), *np.percentile(steps, [50, 90, 99, 99.9, 99.99]).astype(int), steps.max()))
print("tris/ray max %d, nodes/ray max %d" % (c[:, 1].max(), c[:, 0].max()))
heavy = np.argsort(-steps)[:32]
print("heaviest rays (index, row, steps):", [(int(i), int(i) // 1920, int(steps[i])) for i in heavy[:8]])
print("share of all steps in the top 0.1%% rays: %.2f%%" % (100 * np.sort(steps)[-n // 1000:].sum() / steps.sum()))

t = torch.empty(n, device="cuda"); pid = torch.empty(n, dtype=torch.int32, device="cuda")
uv = torch.empty(n, 2, device="cuda"); p = torch.empty(n, 3, device="cuda")
valid = torch.empty(n, dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(ptr, m, reps=9):
    ms = []
    for _ in range(reps):
        flush.zero_()
        torch.cuda.synchronize()
        ms.append(b.trace_dev(ptr, m, hitreg.ALL, uv.data_ptr(), t.data_ptr(), pid.data_ptr(),
                              p.data_ptr(), valid.data_ptr()))
    return float(np.median(ms[2:]))


print("as is                      %.4f ms" % run(d_rays.data_ptr(), n))
# heavy rays first, rest in original order
k = n // 200
order = np.concatenate([np.sort(heavy_all := np.argsort(-steps)[:k]), np.setdiff1d(np.arange(n), heavy_all)])
d2 = torch.from_numpy(rays[order]).cuda()
print("top 0.5%% rays first         %.4f ms" % run(d2.data_ptr(), n))
# only the cheap 99.5 %
d3 = torch.from_numpy(rays[np.setdiff1d(np.arange(n), heavy_all)]).cuda()
print("without the top 0.5%% rays   %.4f ms (%d rays)" % (run(d3.data_ptr(), n - k), n - k))
d4 = torch.from_numpy(rays[np.sort(heavy_all)]).cuda()
print("only the top 0.5%% rays      %.4f ms (%d rays)" % (run(d4.data_ptr(), k), k))
d5 = torch.from_numpy(rays[heavy[:32]]).cuda()
print("only the 32 heaviest rays   %.4f ms" % run(d5.data_ptr(), 32))
d6 = torch.from_numpy(rays[heavy[:1]].repeat(32, 0)).cuda()
print("the heaviest ray x32        %.4f ms" % run(d6.data_ptr(), 32))

print("--- each of the 32 heaviest rays alone (x32 lanes), warm L2")
def run_warm(ptr, m, reps=7):
    ms = []
    for _ in range(reps):
        torch.cuda.synchronize()
        ms.append(b.trace_dev(ptr, m, hitreg.ALL, uv.data_ptr(), t.data_ptr(), pid.data_ptr(),
                              p.data_ptr(), valid.data_ptr()))
    return float(np.median(ms[2:]))
for j in range(32):
    i = int(heavy[j]) if j < len(heavy) else int(np.argsort(-steps)[j])
    dj = torch.from_numpy(rays[i:i + 1].repeat(32, 0)).cuda()
    print("ray %8d (row %4d col %4d) steps %3d  dir %s  %.4f ms" % (
        i, i // 1920, i % 1920, steps[i], np.array2string(rays[i, 3:], precision=6), run_warm(dj.data_ptr(), 32)))
top32 = np.argsort(-steps)[:32]
d7 = torch.from_numpy(rays[top32]).cuda()
print("32 heaviest together, warm L2: %.4f ms" % run_warm(d7.data_ptr(), 32))
for m in (2, 4, 8, 16):
    d8 = torch.from_numpy(rays[top32[:m]]).cuda()
    print("%2d heaviest together, warm L2: %.4f ms" % (m, run_warm(d8.data_ptr(), m)))
