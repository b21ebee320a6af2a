"""Multi-GPU plumbing: rays shard, the scene replicates.

One process per GPU over torch.distributed (NCCL on the GPUs, gloo in the CPU-only tests).  The
reference has no multi-device path at all (every backend uses device 0); this is the new
functionality `north_star` asks for: the triangle buffer is broadcast, every rank builds the
identical BVH, ray i of R goes to rank floor(i*G/R) (contiguous slices whose sizes differ by at
most one), each rank traces its slice and the hit records come back to rank 0 in ray order.
There is no collective inside the traversal itself.
"""
from __future__ import annotations

import numpy as np


def slice_bounds(n: int, world: int, rank: int):
    """[lo, hi) of rank's contiguous slice: ray i belongs to rank floor(i*world/n)."""
    lo = -(-rank * n // world)        # ceil(rank*n/world)
    hi = -(-(rank + 1) * n // world)
    return lo, hi


def owner_of(i: int, n: int, world: int) -> int:
    return i * world // n


def _dist():
    import torch.distributed as dist
    return dist


def bind_near_gpu(device_index: int):
    """Pin this process (and the threads it starts later) to the CPUs NVML reports as local to
    the GPU, so that the pages this rank touches first -- its slice of a SharedHostBatch -- are
    allocated on the NUMA node its PCIe link hangs off.  Without it a host batch written by one
    process sits on one socket and the GPUs of the other socket DMA across the inter-socket link.
    Placement only: returns the CPU list, or None when NVML or the cpuset do not allow it."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:  # NVML numbers the physical devices
            device_index = int(vis.split(",")[device_index])
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n_words = (os.cpu_count() + 63) // 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, x in enumerate(words) for b in range(64) if (int(x) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


def broadcast_scene(tris, device, src: int = 0, group=None):
    """Rank `src` passes its (N,9) float32 triangles, the others pass None; everyone gets a tensor
    on `device` (36*N bytes over NVLink with NCCL)."""
    import torch
    dist = _dist()
    rank = dist.get_rank(group)
    n = torch.tensor([0 if tris is None else len(tris)], dtype=torch.int64, device=device)
    dist.broadcast(n, src=src, group=group)
    if rank == src:
        t = torch.as_tensor(np.ascontiguousarray(tris, np.float32)).reshape(-1, 9).to(device)
    else:
        t = torch.empty((int(n.item()), 9), dtype=torch.float32, device=device)
    if t.numel():
        dist.broadcast(t, src=src, group=group)
    return t


def scatter_rays(rays, device, src: int = 0, group=None):
    """Rank `src` passes all (R,6) rays; every rank gets (its slice on `device`, R)."""
    import torch
    dist = _dist()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n_here = 0 if rays is None else (rays.numel() // 6 if isinstance(rays, torch.Tensor) else len(rays))
    n = torch.tensor([n_here], dtype=torch.int64, device=device)
    dist.broadcast(n, src=src, group=group)
    R = int(n.item())
    width = max(1, -(-R // world))  # equal padded slices for the collective
    out = torch.zeros((width, 6), dtype=torch.float32, device=device)
    parts = None
    if rank == src:
        if isinstance(rays, torch.Tensor):  # e.g. a pinned host tensor: one DMA, no staging
            full = rays.reshape(-1, 6).to(device, non_blocking=True)
        else:
            full = torch.as_tensor(np.ascontiguousarray(rays, np.float32)).reshape(-1, 6).to(device)
        parts = []
        for r in range(world):
            lo, hi = slice_bounds(R, world, r)
            if hi - lo == width:
                parts.append(full[lo:hi])  # equal slices: views, no copy
            else:
                p = torch.zeros((width, 6), dtype=torch.float32, device=device)
                p[: hi - lo] = full[lo:hi]
                parts.append(p)
    dist.scatter(out, parts, src=src, group=group)
    lo, hi = slice_bounds(R, world, rank)
    return out[: hi - lo].contiguous(), R


def gather_device(local, R: int, dst: int = 0, group=None):
    """local: uint8 tensor (n_local, stride) of this rank's hit records, on the collective's
    device.  Rank `dst` gets a (R, stride) uint8 tensor in ray order, the others None."""
    import torch
    dist = _dist()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    stride = local.shape[1]
    width = max(1, -(-R // world))
    if local.shape[0] == width:
        buf = local.contiguous()
    else:
        buf = torch.zeros((width, stride), dtype=torch.uint8, device=local.device)
        buf[: local.shape[0]] = local
    if rank != dst:
        dist.gather(buf, None, dst=dst, group=group)
        return None
    out = torch.empty((R, stride), dtype=torch.uint8, device=local.device)
    if R == width * world:  # equal slices: receive straight into the ray-ordered result
        parts = [out[r * width:(r + 1) * width] for r in range(world)]
        dist.gather(buf, parts, dst=dst, group=group)
        return out
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.gather(buf, parts, dst=dst, group=group)
    for r in range(world):
        lo, hi = slice_bounds(R, world, r)
        if hi > lo:
            out[lo:hi] = parts[r][: hi - lo]
    return out


def gather_hits(local_hits: np.ndarray, R: int, device, dst: int = 0, group=None):
    """local_hits: structured HitReg array of this rank's slice.  Rank `dst` gets the (R,) array in
    ray order, the others None."""
    import torch
    stride = local_hits.dtype.itemsize
    raw = np.ascontiguousarray(local_hits).view(np.uint8).reshape(len(local_hits), stride)
    out = gather_device(torch.from_numpy(raw).to(device), R, dst, group)
    if out is None:
        return None
    return out.cpu().numpy().reshape(-1).view(local_hits.dtype).copy()


class SharedHostBatch:
    """Single-node fast path for host-resident batches: the ray batch and the hit records live in
    one POSIX shared-memory segment each (/dev/shm); every rank page-locks ITS OWN contiguous slice
    and lets the host entry point DMA it straight to/from its GPU.  All PCIe links work in parallel
    and the hits land in ray order in the shared result -- no funnel through rank 0, no collective
    on the data path.  (The NCCL scatter/gather above is for batches that live on a GPU.)"""

    def __init__(self, name: str, n_rays: int, hit_dtype, rank: int, world: int, create: bool,
                 first_touch: bool = False):
        """first_touch=True zero-fills this rank's slices so that their pages are allocated on the
        NUMA node the rank runs on (see bind_near_gpu).  Only for a batch nobody has written yet:
        construct on every rank, synchronise, then fill."""
        import ctypes as C

        from ._lib import lib
        self.n, self.rank, self.world = n_rays, rank, world
        self.paths = (f"/dev/shm/{name}_rays", f"/dev/shm/{name}_hits")
        mode = "w+" if create else "r+"
        self.rays = np.memmap(self.paths[0], np.float32, mode, shape=(n_rays, 6))
        self.hits = np.memmap(self.paths[1], np.dtype(hit_dtype), mode, shape=(n_rays,))
        self.lo, self.hi = slice_bounds(n_rays, world, rank)
        if first_touch:
            self.rays[self.lo:self.hi] = 0
            self.hits[self.lo:self.hi] = np.zeros((), np.dtype(hit_dtype))
        self._reg = []
        page = 4096
        for arr, item in ((self.rays, 24), (self.hits, np.dtype(hit_dtype).itemsize)):
            base = arr.ctypes.data
            a = (base + self.lo * item) // page * page
            b = -(-(base + self.hi * item) // page) * page
            end = -(-(base + n_rays * item) // page) * page
            b = min(b, end)
            if b > a and lib().prt_b200_host_register(C.c_void_p(a), b - a) == 0:
                self._reg.append(a)

    @property
    def my_rays(self):
        return self.rays[self.lo:self.hi]

    @property
    def my_hits(self):
        return self.hits[self.lo:self.hi]

    def close(self, unlink: bool):
        import ctypes as C
        import os

        from ._lib import lib
        for a in self._reg:
            lib().prt_b200_host_unregister(C.c_void_p(a))
        self._reg = []
        if unlink:
            for p in self.paths:
                try:
                    os.unlink(p)
                except OSError:
                    pass


def sharded_nearest_hits(set_tris, trace, tris, rays, device, src: int = 0, group=None):
    """The whole multi-GPU call: `set_tris(tris_np)` and `trace(rays_np) -> structured hits` are the
    per-rank backend calls (CUDABackend.set_tris / .nearest_hits bound to this rank's GPU); `tris`
    and `rays` are given on rank `src` only.  Returns the ray-ordered hits on rank `src`."""
    t = broadcast_scene(tris, device, src, group)
    set_tris(t.cpu().numpy())
    mine, R = scatter_rays(rays, device, src, group)
    hits = trace(mine.cpu().numpy())
    return gather_hits(hits, R, device, src, group)
