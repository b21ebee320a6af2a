"""Experiment: how much does ray reordering buy on incoherent batches?  Sort on the host (numpy),
trace on the device, compare device time with the unsorted batch."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import portablert_b200 as prt
from portablert_b200 import scenes, hitreg

def morton3(q):  # q: (n,3) ints < 1024
    def spread(v):
        v = v.astype(np.uint64) & 0x3ff
        v = (v | (v << 16)) & 0x30000ff
        v = (v | (v << 8)) & 0x300f00f
        v = (v | (v << 4)) & 0x30c30c3
        v = (v | (v << 2)) & 0x9249249
        return v
    return (spread(q[:,0]) << 2) | (spread(q[:,1]) << 1) | spread(q[:,2])

def keys(rays, lo, hi, mode):
    o = rays[:, :3]; d = rays[:, 3:6]
    qo = np.clip(((o - lo) / (hi - lo) * 1024).astype(np.int64), 0, 1023)
    dn = d / np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-30)
    octant = ((dn[:,0] < 0).astype(np.uint64) << 2) | ((dn[:,1] < 0).astype(np.uint64) << 1) | (dn[:,2] < 0).astype(np.uint64)
    qd = np.clip(((dn * 0.5 + 0.5) * 16).astype(np.int64), 0, 15)
    if mode == "origin":
        return morton3(qo)
    if mode == "octant_origin":
        return (octant << 30) | morton3(qo)
    if mode == "dir_origin":   # 12 direction bits on top, then origin
        return (morton3(qd) << 30) | morton3(qo)
    if mode == "origin_dir":   # coarse origin (5 bits/axis), then direction, then fine origin
        return (morton3(qo >> 5) << 27) | (morton3(qd) << 15) | morton3(qo & 31)
    if mode == "o12_d12":      # 24 bits: 3 passes
        return (morton3(qo >> 6) << 12) | morton3(qd)
    if mode == "o15_d12":      # 27 bits
        return (morton3(qo >> 5) << 12) | morton3(qd)
    if mode == "o12_d12_o8":   # 32 bits: 4 passes
        return (morton3(qo >> 6) << 20) | (morton3(qd) << 8) | (morton3((qo >> 3) & 7) >> 1)
    if mode == "o15_d9":       # 24 bits, coarser direction (3 bits/axis)
        return (morton3(qo >> 5) << 9) | morton3(qd >> 1)
    raise ValueError(mode)

def run(name, tris, rays, mask):
    b = prt.cuda_backend
    dev = torch.device("cuda", 0)
    d_tris = torch.from_numpy(tris).to(dev)
    torch.cuda.synchronize()
    b.set_tris_dev(d_tris.data_ptr(), len(tris))
    lo = tris.reshape(-1,3).min(0); hi = tris.reshape(-1,3).max(0)
    n = len(rays)
    t = torch.empty(n, device=dev); pid = torch.empty(n, dtype=torch.int32, device=dev)
    uv = torch.empty(n, 2, device=dev); p = torch.empty(n, 3, device=dev); valid = torch.empty(n, dtype=torch.uint8, device=dev)
    def time_it(r):
        d_rays = torch.from_numpy(np.ascontiguousarray(r)).to(dev)
        torch.cuda.synchronize()
        ms = []
        for k in range(6):
            ms.append(b.trace_dev(d_rays.data_ptr(), n, mask, uv=uv.data_ptr(), t=t.data_ptr(), pid=pid.data_ptr(), p=p.data_ptr(), valid=valid.data_ptr()))
        return float(np.mean(ms[2:]))
    base = time_it(rays)
    print(f"{name}: unsorted {base:.3f} ms ({n/base/1e3:.0f} Mrays/s)")
    for mode in ("origin_dir", "o12_d12", "o15_d12", "o12_d12_o8", "o15_d9"):
        k = keys(rays, lo, hi, mode)
        order = np.argsort(k, kind="stable")
        ms = time_it(rays[order])
        print(f"   sorted by {mode:14s} {ms:.3f} ms ({n/ms/1e3:.0f} Mrays/s)  x{base/ms:.2f}")

prt.select_backend(prt.cuda_backend)
tris = scenes.sphere_field(2000)
lo, hi = tris.reshape(-1,3).min(0), tris.reshape(-1,3).max(0)
tris = scenes.sphere_field(10000)
lo, hi = tris.reshape(-1,3).min(0), tris.reshape(-1,3).max(0)
run("C4 (10M tris, 10M rays)", tris, scenes.incoherent_rays(10_000_000, lo, hi, 4), hitreg.T | hitreg.PID)
tris = scenes.interior()
prim = scenes.camera_rays(3840, 2160, (2, 6, 3), (28, 4, 15))
prt.cuda_backend.set_tris(tris)
h = prt.cuda_backend.nearest_hits(prim)
pp = np.stack([h["px"], h["py"], h["pz"]], -1)
b_rays, _ = scenes.bounce_rays(tris, prim, h["valid"], h["primitive_id"], pp)
run("C3 bounce", tris, b_rays, hitreg.T | hitreg.PID)
