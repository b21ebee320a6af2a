import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import portablert_b200 as prt
from portablert_b200 import scenes, hitreg
prt.select_backend(prt.cuda_backend); b = prt.cuda_backend
tris = scenes.blob(); dev = torch.device("cuda", 0)
d_tris = torch.from_numpy(tris).to(dev); torch.cuda.synchronize()
b.set_tris_dev(d_tris.data_ptr(), len(tris))
for frame in range(4):
    rays = scenes.pinhole_rays(1920, 1080, cam=(0.002 * frame, 0.0, -0.3))
    n = len(rays); d_rays = torch.from_numpy(rays).to(dev)
    uv = torch.empty(n, 2, device=dev); t = torch.empty(n, device=dev); pid = torch.empty(n, dtype=torch.int32, device=dev)
    p = torch.empty(n, 3, device=dev); valid = torch.empty(n, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    ms = [b.trace_dev(d_rays.data_ptr(), n, 31, uv.data_ptr(), t.data_ptr(), pid.data_ptr(), p.data_ptr(), valid.data_ptr()) for _ in range(12)]
    zero = (rays[:, 3:6] == 0).any(1).sum()
    print(f"frame {frame}: {np.mean(ms[4:]):.4f} ms, valid {valid.float().mean().item():.4f}, rays with a zero dir component {zero}")
    cnt = torch.zeros(n, 2, dtype=torch.int32, device=dev); torch.cuda.synchronize()
    b.trace_count_dev(d_rays.data_ptr(), n, cnt.data_ptr())
    c = cnt.to(torch.float64)
    print(f"   nodes/ray {c[:,0].mean().item():.3f} tris/ray {c[:,1].mean().item():.3f} max nodes {int(cnt[:,0].max())} max tris {int(cnt[:,1].max())}; rays>100 nodes: {(cnt[:,0]>100).sum().item()}, rays>50 tris: {(cnt[:,1]>50).sum().item()}")
