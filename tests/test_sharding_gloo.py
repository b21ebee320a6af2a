"""World-size-2/3 gloo tests (CPU) of the multi-GPU plumbing: broadcast the scene, slice the rays
contiguously, trace per rank, gather in ray order.  Invariant: the N-rank result is
byte-for-byte the 1-rank result.  The per-rank tracer here is the oracle (no GPU in this suite);
on the GPU box bench.py runs the same plumbing over NCCL with the CUDA backend."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slice_bounds_partition():
    from portablert_b200.sharding import owner_of, slice_bounds
    for n in (0, 1, 2, 7, 31, 32, 33, 1000, 2_073_600):
        for world in (1, 2, 3, 4, 8):
            b = [slice_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
            for i in (0, n // 3, n - 1):
                if 0 <= i < n:
                    r = owner_of(i, n, world)
                    assert b[r][0] <= i < b[r][1]


def _worker(rank, world, port, n_rays, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    from oracle import Oracle
    from portablert_b200 import hitreg, scenes, sharding
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank,
                            world_size=world)
    orc = Oracle()
    dt = hitreg.dtype(hitreg.ALL)

    def trace(r):
        o = orc.trace(r, threads=1)
        h = np.zeros(len(r), dt)
        for src, dst in (("t", "t"), ("u", "u"), ("v", "v"), ("pid", "primitive_id"),
                         ("valid", "valid"), ("px", "px"), ("py", "py"), ("pz", "pz")):
            h[dst] = o[src]
        return h

    tris = scenes.blob(20, 20) if rank == 0 else None
    rays = scenes.pinhole_rays(n_rays // 10 + 1, 10)[:n_rays] if rank == 0 else None
    hits = sharding.sharded_nearest_hits(orc.build, trace, tris, rays, torch.device("cpu"))
    if rank == 0:
        np.save(out_path, hits)
    else:
        assert hits is None
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,n_rays", [(2, 1000), (3, 1001), (2, 1)])
def test_n_ranks_equal_one_rank(tmp_path, world, n_rays):
    outs = []
    for w in (1, world):
        out = str(tmp_path / f"hits_{w}.npy")
        mp.spawn(_worker, args=(w, _free_port(), n_rays, out), nprocs=w, join=True)
        outs.append(np.load(out))
    assert len(outs[0]) == len(outs[1]) == n_rays
    for f in outs[0].dtype.names:  # field-wise: numpy does not preserve struct padding bytes
        a, b = np.ascontiguousarray(outs[0][f]), np.ascontiguousarray(outs[1][f])
        assert a.view(np.uint8).tobytes() == b.view(np.uint8).tobytes(), f


def _shm_worker(rank, world, n_rays, tag):
    sys.path.insert(0, ROOT)
    from oracle import Oracle
    from portablert_b200 import hitreg, scenes, sharding
    import time
    dt = hitreg.dtype(hitreg.T | hitreg.VALID)
    if rank == 0:
        shm = sharding.SharedHostBatch(tag, n_rays, dt, rank, world, True)
        shm.rays[...] = scenes.pinhole_rays(n_rays // 10 + 1, 10)[:n_rays]
        shm.rays.flush()
        open(f"/dev/shm/{tag}_ready", "w").close()
    else:
        while not os.path.exists(f"/dev/shm/{tag}_ready"):
            time.sleep(0.01)
        shm = sharding.SharedHostBatch(tag, n_rays, dt, rank, world, False)
    orc = Oracle().build(scenes.blob(20, 20))
    o = orc.trace(np.asarray(shm.my_rays), threads=1)
    mine = shm.my_hits
    mine["t"] = o["t"]
    mine["valid"] = o["valid"]
    shm.hits.flush()
    shm.close(unlink=False)


def test_shared_host_batch_slices_assemble_in_ray_order():
    """The multi-GPU host layout without GPUs: every rank fills only its own slice of the shared
    result; together they must be exactly the single-process answer, in ray order."""
    from oracle import Oracle
    from portablert_b200 import hitreg, scenes
    n_rays, world = 1003, 3
    tag = "prt_b200_gloo_%d" % os.getpid()
    try:
        mp.spawn(_shm_worker, args=(world, n_rays, tag), nprocs=world, join=True)
        dt = hitreg.dtype(hitreg.T | hitreg.VALID)
        hits = np.memmap(f"/dev/shm/{tag}_hits", dt, "r", shape=(n_rays,))
        rays = scenes.pinhole_rays(n_rays // 10 + 1, 10)[:n_rays]
        ref = Oracle().build(scenes.blob(20, 20)).trace(rays)
        assert np.array_equal(np.asarray(hits["t"]), ref["t"])
        assert np.array_equal(np.asarray(hits["valid"]), ref["valid"])
    finally:
        for suffix in ("_rays", "_hits", "_ready"):
            try:
                os.unlink(f"/dev/shm/{tag}{suffix}")
            except OSError:
                pass
