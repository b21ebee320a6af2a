"""Worker of tests/test_gpu_scale.py::test_process_per_gpu_sharding_over_nccl (run under torchrun,
one process per GPU): portablert_b200.sharding over NCCL on real GPUs vs the single-GPU result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import portablert_b200 as prt  # noqa: E402
from portablert_b200 import scenes, sharding  # noqa: E402


def main():
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    b = prt.CUDABackend(device=local)
    b.init()
    tris = scenes.interior(30_000) if rank == 0 else None
    lo, hi = (0, 0, 0), (30, 12, 18)
    rays = np.concatenate([scenes.camera_rays(320, 181, (2, 6, 3), (28, 4, 15)),
                           scenes.incoherent_rays(100_003, lo, hi, seed=5)]) if rank == 0 else None
    hits = sharding.sharded_nearest_hits(b.set_tris, lambda r: b.nearest_hits(r), tris, rays, dev)
    ok = torch.ones(1, device=dev)
    if rank == 0:
        b.set_tris(tris)
        one = b.nearest_hits(rays)
        same = all(np.array_equal(one[f], hits[f], equal_nan=True) for f in one.dtype.names)
        ok[0] = 1.0 if (same and len(hits) == len(rays)) else 0.0
    dist.broadcast(ok, src=0)
    dist.barrier()
    if rank == 0:
        print("SHARDING_OK" if ok.item() == 1.0 else "SHARDING_MISMATCH", world, "ranks,", len(rays), "rays")
    b.shutdown()
    dist.destroy_process_group()
    sys.exit(0 if ok.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
