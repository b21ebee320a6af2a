"""-m gpu: the paths only large scenes, unusual rays or several GPUs reach -- 48/63-bit Morton keys
with 6 and 8 sort passes, the compressed wide nodes on a >= 2^20-triangle scene, the stack overflow
area, the exact second pass for rays the fast box test cannot take, packed result transfers, and
the multi-GPU context of the library (G GPUs == 1 GPU byte for byte)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import parity
import portablert_b200 as prt
from portablert_b200 import hitreg, scenes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def soa(h):
    return parity.from_structured(h)


def n_gpus():
    import torch
    return torch.cuda.device_count()


def same_hits(a, b):
    """field-wise equality (padding / Empty bytes of a record are unspecified, as in the reference)"""
    return a.dtype == b.dtype and len(a) == len(b) and all(
        np.array_equal(a[f], b[f], equal_nan=True) for f in a.dtype.names)


def fresh_backend(env=None, **kw):
    """A second backend object beside the session's, created under extra environment knobs (they
    are read in prt_b200_create)."""
    old = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update(env or {})
    try:
        b = prt.CUDABackend(**kw)
        b.init()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return b


def test_million_triangle_scene_key_widths_and_wide_nodes(oracle):
    """>= 2^20 triangles: the default picks 13 bits per axis here, C4 (10 M) picks 16; both, and
    the 21-bit / 8-pass maximum, are forced on the same scene.  The incoherent batch is reordered
    and therefore traced through the compressed 4-wide nodes (mode 2); every variant must
    reproduce the oracle, and all variants each other byte for byte."""
    tris = scenes.sphere_field(1100, extent=300.0)  # 1 100 000 triangles
    assert len(tris) >= 1 << 20
    lo, hi = tris.reshape(-1, 3).min(0), tris.reshape(-1, 3).max(0)
    rays = scenes.incoherent_rays(400_000, lo, hi, seed=9)
    oracle.build(tris)
    sel = np.arange(0, len(rays), 4)
    ref = oracle.trace(rays[sel])
    want = None
    for bits in ("13", "16", "21"):
        b = fresh_backend({"PRT_B200_MORTON_BITS": bits})
        try:
            b.set_tris(tris)
            before = b.sorted_batches
            h = b.nearest_hits(rays, "t", "primitive_id")
            assert b.sorted_batches > before, "incoherent batch must be reordered (-> wide nodes)"
            got = {"t": h["t"][sel], "pid": h["primitive_id"][sel], "valid": np.isfinite(h["t"][sel])}
            rep = parity.compare({k: ref[k] for k in ("t", "pid", "valid")}, got, tris, rays[sel], oracle)
            parity.assert_parity(rep)
            assert rep["t_equal"], (bits, rep)
            # binary nodes on the same batch: identical records
            b.set_wide_nodes(0)
            b.set_tris(tris)
            h2 = b.nearest_hits(rays, "t", "primitive_id")
            assert same_hits(h, h2), bits
            if want is None:
                want = h.copy()
            assert same_hits(h, want), f"{bits}-bit keys changed a result"
        finally:
            b.shutdown()


def test_deep_stacks_overflow_into_global_memory(cuda, oracle):
    """131 072 copies of one triangle: equal Morton keys, the radix tree splits on the index and
    is 17 levels deep, and a ray through the triangle finds BOTH children of every node -- the
    traversal stack outgrows its shared-memory part.  All hits tie; lowest primitive id wins."""
    tri = np.array([[-1, -1, 0, 1, -1, 0, 0, 1, 0]], np.float32)
    tris = np.repeat(tri, 1 << 17, axis=0)
    rays = scenes.c1_rays(4096, seed=3)
    cuda.set_tris(tris)
    for prune in (0, 1):
        cuda.set_trace_opts(prune=prune)
        h = cuda.nearest_hits(rays)
        cuda.set_trace_opts()
        ref = oracle.brute(tri, rays)
        assert np.array_equal(h["valid"], ref["valid"])
        assert np.array_equal(h["t"], ref["t"])
        assert (h["primitive_id"][h["valid"]] == 0).all()


def weird_rays(lo, hi, n=4000, seed=5):
    """Axis-parallel rays (one and two zero components, both signs), rays whose origin lies exactly
    on box face planes, denormal / huge / non-finite components."""
    g = np.random.default_rng(seed)
    r = scenes.incoherent_rays(n, lo, hi, seed)
    k = n // 8
    r[0 * k:1 * k, 3] = 0.0
    r[1 * k:2 * k, [3, 4]] = 0.0
    r[2 * k:3 * k, 5] = -0.0
    r[3 * k:4 * k, 4] = 1e-30
    r[4 * k:5 * k, 3:6] *= np.float32(1e-18)
    r[5 * k:6 * k, 0:3] *= np.float32(1e6)
    bad = r[6 * k:7 * k]
    bad[0::5, 0] = np.nan
    bad[1::5, 4] = np.inf
    bad[2::5, 3:6] = 0.0
    bad[3::5, 1] = np.float32(3e38)
    bad[4::5, 5] = np.float32(1e-42)  # denormal: 1/d overflows
    # origins snapped onto vertex coordinates of the scene (face planes of leaf boxes)
    return r


def test_axis_parallel_and_exotic_rays(cuda, oracle):
    tris = scenes.blob(48, 48)
    lo, hi = tris.reshape(-1, 3).min(0), tris.reshape(-1, 3).max(0)
    rays = weird_rays(lo - 0.05, hi + 0.05)
    snap = rays[: len(rays) // 8]
    verts = tris.reshape(-1, 3)
    snap[:, 0] = verts[np.arange(len(snap)) * 7 % len(verts), 0]  # d_x == 0 and o_x on a face plane
    oracle.build(tris)
    cuda.set_tris(tris)
    import torch
    dev = torch.device("cuda", 0)
    d_rays = torch.from_numpy(rays).to(dev)
    t = torch.zeros(len(rays), device=dev)
    pid = torch.zeros(len(rays), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    before = cuda.exotic_rays
    cuda.trace_dev(d_rays.data_ptr(), len(rays), 6, t=t.data_ptr(), pid=pid.data_ptr())
    assert cuda.exotic_rays > before, "non-finite rays must reach the exact second pass"
    assert cuda.exotic_rays - before <= len(rays) // 4, "axis-parallel rays must stay on the fast path"
    host = cuda.nearest_hits(rays, "t", "primitive_id")  # (the host pipeline launches the pass inline)
    assert np.array_equal(t.cpu().numpy(), host["t"], equal_nan=True)
    exact = fresh_backend({"PRT_B200_FAST_BOXES": "0"})
    try:
        exact.set_tris(tris)
        for prune in (1, 0):
            cuda.set_trace_opts(prune=prune)
            exact.set_trace_opts(prune=prune)
            a, b = cuda.nearest_hits(rays), exact.nearest_hits(rays)
            for f in a.dtype.names:
                assert np.array_equal(a[f], b[f], equal_nan=True), (prune, f)
        cuda.set_trace_opts()
    finally:
        exact.shutdown()
    # against the reference's per-triangle rule (brute force: no topology involved)
    fin = np.isfinite(rays).all(1)
    ref = oracle.brute(tris, rays[fin])
    got = soa(cuda.nearest_hits(rays[fin]))
    rep = parity.compare(ref, got, tris, rays[fin], oracle)
    parity.assert_parity(rep)
    assert rep["t_equal"], rep


def test_packed_results_equal_direct_dma(cuda):
    """Pageable results cross PCIe tightly packed and are scattered into the caller's records;
    pinned results are DMA'd as records.  Field for field the same, for all 31 layouts, and fewer
    bytes moved where the record has padding."""
    from portablert_b200.backend import pinned_empty
    tris = scenes.blob(40, 40)
    rays = np.concatenate([scenes.pinhole_rays(400, 300), scenes.c1_rays(50_001, seed=2) * 0.1])
    cuda.set_tris(tris)
    p_rays = pinned_empty(rays.shape, np.float32)
    p_rays[...] = rays
    for combo in hitreg.TAG_COMBOS:
        m = hitreg.mask_of(combo)
        pin = pinned_empty((len(rays),), hitreg.dtype(m))
        cuda.nearest_hits(p_rays, m, out=pin)
        _, d2h_pin = cuda.last_transfer_bytes
        page = cuda.nearest_hits(rays, m)
        _, d2h_page = cuda.last_transfer_bytes
        for f in page.dtype.names:
            assert np.array_equal(page[f], pin[f], equal_nan=True), (combo, f)
        assert d2h_pin == len(rays) * hitreg.layout(m)[0] and d2h_page <= d2h_pin
    page = cuda.nearest_hits(rays, "t", "primitive_id")
    assert cuda.last_transfer_bytes[1] == 8 * len(rays)  # 16-byte records, 8 bytes on the wire
    page = cuda.nearest_hits(rays, "valid")
    assert cuda.last_transfer_bytes[1] == len(rays)


@pytest.mark.parametrize("gpus", [2, 4, 8])
def test_multi_gpu_context_equals_one_gpu(cuda, gpus):
    """prt_b200_create_multi: the scene is broadcast over NVLink and built on every device, a host
    batch is cut into contiguous slices and the hits land in ray order.  Must equal the single-GPU
    result byte for byte, for pageable and pinned buffers, coherent and reordered batches, and
    ragged batch sizes."""
    if n_gpus() < gpus:
        pytest.skip(f"needs {gpus} GPUs")
    from portablert_b200.backend import pinned_empty
    tris = scenes.interior(60_000)
    lo, hi = tris.reshape(-1, 3).min(0), tris.reshape(-1, 3).max(0)
    batches = [scenes.camera_rays(640, 363, (2, 6, 3), (28, 4, 15)),
               scenes.incoherent_rays(700_001, lo, hi, seed=21),
               scenes.incoherent_rays(5, lo, hi, seed=22)]  # fewer rays than GPUs
    cuda.set_tris(tris)
    for path in ("nccl", "p2p"):
        mg = fresh_backend({"PRT_B200_BCAST": path}, gpus=gpus)
        try:
            assert mg.num_devices == gpus
            mg.set_tris(tris)
            assert mg.broadcast_path in ("nccl", "p2p") and (path == "nccl" or mg.broadcast_path == "p2p")
            print(f"{gpus} GPUs, triangles broadcast via {mg.broadcast_path}: {mg.device_name()}")
            for rays in batches:
                for combo in (hitreg.TAG_COMBOS[-1], ("t", "primitive_id"), ("valid",)):
                    one = cuda.nearest_hits(rays, *combo)
                    many = mg.nearest_hits(rays, *combo)
                    assert same_hits(one, many), (path, len(rays), combo)
            rays = batches[1]
            p_rays = pinned_empty(rays.shape, np.float32)
            p_rays[...] = rays
            p_hits = pinned_empty((len(rays),), hitreg.dtype(hitreg.ALL))
            mg.nearest_hits(p_rays, hitreg.ALL, out=p_hits)
            one = cuda.nearest_hits(rays)
            for f in one.dtype.names:
                assert np.array_equal(one[f], p_hits[f], equal_nan=True), f
            # a new scene replaces the replicas everywhere; the empty scene too
            small = scenes.blob(20, 20)
            mg.set_tris(small)
            cuda.set_tris(small)
            r = scenes.pinhole_rays(333, 211)
            assert same_hits(mg.nearest_hits(r), cuda.nearest_hits(r))
            mg.set_tris(np.zeros((0, 9), np.float32))
            assert not mg.nearest_hits(r, "valid")["valid"].any()
            cuda.set_tris(tris)
        finally:
            mg.shutdown()


def test_dropin_program_spreads_over_all_gpus():
    """The reference's own C++ API with this backend, unchanged source, PRT_B200_GPUS = all GPUs:
    select_backend / set_tris / nearest_hits<Tags...> for the 31 tag combinations vs the reference's
    CPU backend in the same process."""
    if n_gpus() < 2:
        pytest.skip("needs >= 2 GPUs")
    exe = os.path.join(ROOT, "oracle", "_ref", "dropin_test")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/dropin_test not built")
    env = dict(os.environ, PRT_B200_GPUS=str(n_gpus()))
    out = subprocess.run([exe], capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert f"x{n_gpus()}" in out.stdout, out.stdout  # device_name() reports the GPU count
    print(out.stdout[-800:])


def test_process_per_gpu_sharding_over_nccl():
    """portablert_b200.sharding (one process per GPU, torch.distributed): NCCL broadcast of the
    scene, NCCL scatter of the rays, NCCL gather of the hit records -- on real GPUs, against the
    single-GPU result."""
    if n_gpus() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(4, n_gpus())
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29731",
           os.path.join(ROOT, "tests", "nccl_sharding_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "SHARDING_OK" in out.stdout, out.stdout[-2000:]


def test_rebuild_chain_is_replayed_from_a_cuda_graph(cuda):
    """Small scenes: the ~11 launches of a rebuild are captured into a CUDA graph at the second build
    from the same buffers and replayed afterwards.  Every build must produce the same BVH bytes, a
    different scene in between must not be served from the stale graph, and tracing still works."""
    import torch
    dev = torch.device("cuda", 0)
    a = torch.from_numpy(scenes.blob(96, 96)).to(dev)
    b_host = scenes.blob(96, 96, radius=0.13, bump=0.3)
    b = torch.from_numpy(b_host).to(dev)
    rays = scenes.pinhole_rays(320, 200)
    torch.cuda.synchronize()
    cuda.set_tree_optimisation(0)
    before = cuda.graph_replays
    want = None
    for k in range(5):
        cuda.set_tris_dev(a.data_ptr(), len(a))
        nodes, recs = cuda.download_bvh()
        if want is None:
            want = (nodes.tobytes(), recs.tobytes(), cuda.nearest_hits(rays).copy())
        assert nodes.tobytes() == want[0] and recs.tobytes() == want[1], k
    assert cuda.graph_replays >= before + 2, "builds 3.. must be graph replays"
    cuda.set_tris_dev(b.data_ptr(), len(b))  # same size, other buffer: not the captured chain
    other = cuda.nearest_hits(rays)
    fresh = fresh_backend({"PRT_B200_GRAPHS": "0"})
    try:
        fresh.set_tree_optimisation(0)
        fresh.set_tris(b_host)
        assert same_hits(other, fresh.nearest_hits(rays))
    finally:
        fresh.shutdown()
    cuda.set_tris_dev(a.data_ptr(), len(a))
    assert same_hits(cuda.nearest_hits(rays), want[2])
    cuda.set_tree_optimisation(3)
