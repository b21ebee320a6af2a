#!/usr/bin/env python
"""bench.py -- device-timed Mrays/s of nearest_hits (+ BVH build Mtris/s) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2|c3|c4|c5] [--impl reference]

One JSON line on stdout (rank 0).  A "step" is one nearest_hits pass over the whole ray batch of
the workload.  Default workload = BASELINE.json configs[1] (C2): synthetic 69 192-triangle
bunny-scale mesh, 1920x1080 coherent primary rays, all five filter tags (FullHitReg).

  value       whole-job Mrays/s, rays resident in HBM, SoA outputs in HBM; per-step device time
              from CUDA events on the stream the traversal kernel runs on (taken by the library
              around the launch), L2 flushed between steps, max over ranks
  e2e         the same metric through the reference-facing host API (host ray buffer in, host
              HitReg records out: H2D + kernel + D2H inside the timed region)
  roofline    the traversal kernel against the measured HBM peak (and the L2 read bandwidth
              measured on the box), from ALGORITHMIC bytes per ray counted by the instrumented
              kernel: 24 (ray) + 64 * nodes fetched + 64 * triangles tested + output bytes
  cpu_baseline  the unmodified reference CPU backend (oracle/_ref) on the box's host cores
  --impl reference  times only that CPU reference and prints the same line shape

N > 1 (torchrun, one rank per GPU): rank 0's triangles are NCCL-broadcast, every rank builds the
identical BVH, the ray batch is N copies of the single-GPU batch (weak scaling; rank r traces the
r-th contiguous slice, no collective on the data path).  In the e2e leg the whole batch lives in
host shared memory, every rank DMA's its own page-locked slice through the host entry point and the
hits land in ray order in the shared result.

Other configs: --config c3 (262 k-tri interior, 4K primary), c3b (its one-bounce diffuse rays),
c5 (1 M-tri height field, 8 M rays), c4 (10 M tris, 100 M incoherent rays; PRT_BENCH_C4_SPHERES /
PRT_BENCH_C4_RAYS scale it down).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# ------------------------------------------------------------------------------------------------
def workload(name: str, frame: int = 0, tracer=None):
    """-> (tris (N,9) f32, rays (R,6) f32, mask, description)"""
    from portablert_b200 import hitreg, scenes
    if name == "c2":
        tris = scenes.blob()
        # frame k > 0 moves the camera slightly so that N-GPU batches are N distinct frames
        rays = scenes.pinhole_rays(1920, 1080, cam=(0.002 * frame, 0.0, -0.3))
        return tris, rays, hitreg.ALL, ("C2: 69192-tri displaced-sphere mesh, 1920x1080 pinhole "
                                        "primary rays, tags uv,t,primitive_id,p,valid")
    if name == "c3":
        tris = scenes.interior()
        rays = scenes.camera_rays(3840, 2160, (2 + 0.01 * frame, 6, 3), (28, 4, 15))
        return tris, rays, hitreg.ALL, ("C3: %d-tri interior, 3840x2160 primary rays, all tags"
                                        % len(tris))
    if name == "c3b":
        # one-bounce incoherent diffuse rays spawned at the primary hits of C3 (the primary hits
        # are input preparation: `tracer(tris, rays) -> full hit records` runs before any timing)
        tris = scenes.interior()
        prim = scenes.camera_rays(3840, 2160, (2 + 0.01 * frame, 6, 3), (28, 4, 15))
        h = tracer(tris, prim)
        p = np.stack([h["px"], h["py"], h["pz"]], -1)
        rays, _ = scenes.bounce_rays(tris, prim, h["valid"], h["primitive_id"], p)
        return tris, rays, hitreg.T | hitreg.PID, ("C3 bounce: %d-tri interior, %d one-bounce "
                                                   "cosine-weighted diffuse rays, t+primitive_id"
                                                   % (len(tris), len(rays)))
    if name == "c5":
        tris = scenes.heightfield(frame)
        rays = scenes.camera_rays(3840, 2160, (10, 6, -4), (10, 0, 5))[:8_000_000]
        return tris, rays, hitreg.T | hitreg.VALID, "C5: 1M-tri heightfield, 8M primary rays, t+valid"
    if name == "c4":
        n_s = int(os.environ.get("PRT_BENCH_C4_SPHERES", "10000"))
        n_r = int(os.environ.get("PRT_BENCH_C4_RAYS", "100000000"))
        tris = scenes.sphere_field(n_s)
        lo, hi = tris.reshape(-1, 3).min(0), tris.reshape(-1, 3).max(0)
        rays = scenes.incoherent_rays(n_r, lo, hi, seed=4 + frame)
        return tris, rays, hitreg.T | hitreg.PID, ("C4: %d-tri sphere field, %d incoherent rays, "
                                                   "t+primitive_id" % (len(tris), n_r))
    raise SystemExit(f"unknown config {name}")


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML DURING the timed region (a thread
    polling every ~2 ms; nvidia-smi's 100 ms floor cannot see a millisecond-scale region)."""

    def __init__(self, index: int):
        self.index = index
        self.sm, self.reasons, self.mx = [], set(), None
        self._stop = threading.Event()
        self.thread = None
        self.err = None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                try:
                    idx = int(vis.split(",")[self.index])
                except Exception:
                    pass
            self.h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.nv = nv
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception as e:  # no NVML: say so in the JSON instead of inventing numbers
            self.err = repr(e)
        return self

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
            except Exception as e:
                self.err = repr(e)
                break
            try:
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception as e:
                self.err = "reasons: " + repr(e)
            time.sleep(0.0005)

    def stop(self):
        self._stop.set()
        if self.thread:
            self.thread.join(timeout=1.0)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.mx, "reasons": ["unavailable: %s" % self.err]}
        out = {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.mx,
               "reasons": sorted(self.reasons), "samples": len(self.sm)}
        if self.err:
            out["sampler_error"] = self.err
        return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


_JSON_OUT = sys.stdout
OUT_BYTES = {1: 8, 2: 4, 4: 4, 8: 12, 16: 1}


def out_bytes(mask):
    return sum(b for bit, b in OUT_BYTES.items() if mask & bit)


# ------------------------------------------------------------------------------------------------
def cpu_reference(tris, rays, mask, budget_s=25.0, steps=1, warmup=0):
    """Time the unmodified reference CPU backend (oracle/_ref; test infrastructure, used here as
    the reported baseline only) on a bounded sample of the workload."""
    from oracle import Reference
    ref = Reference()
    b_s = ref.set_tris(tris)
    # bounded sample: a strided subset keeps the image-space mix of hit and miss rays
    probe = rays[:: max(1, len(rays) // 20000)][:20000]
    ref.nearest_hits(probe, mask, keep=False)
    rate = len(probe) / max(ref.last_trace_s, 1e-6)
    n = int(min(len(rays), max(20000, rate * budget_s / max(1, steps + warmup))))
    stride = max(1, len(rays) // n)
    sample = np.ascontiguousarray(rays[::stride])
    times = []
    for k in range(warmup + steps):
        ref.nearest_hits(sample, mask, keep=False)
        if k >= warmup:
            times.append(ref.last_trace_s)
    s = float(np.mean(times))
    what = ("all %d rays" % len(rays)) if stride == 1 else \
        ("every %d-th ray (%d of %d)" % (stride, len(sample), len(rays)))
    return {
        "value": len(sample) / s / 1e6, "unit": "Mrays/s", "cores": ref.threads, "kind": "reference",
        "sample": f"{what}; set_tris on all {len(tris)} tris (1 thread)",
        "build_mtris_s": len(tris) / b_s / 1e6, "build_s": b_s, "trace_s": s,
        "cpu": ref.device_name(), "n_sample": int(len(sample)),
    }


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    def ref_tracer(t, r):
        from oracle import Reference
        ref = Reference()
        ref.set_tris(t)
        return ref.nearest_hits(r, 31)

    tris, rays, mask, desc = workload(args.config, tracer=ref_tracer)
    cb = cpu_reference(tris, rays, mask, budget_s=60.0, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": "nearest_hits throughput", "value": cb["value"],
        "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": cb["trace_s"] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "tag_mask": mask},
        "build_mtris_s": cb["build_mtris_s"],
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="c2")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--per-mask", action="store_true", help="also time all 31 tag masks")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: keep a private handle to it and send everything else
    # that libraries print to fd 1 (e.g. NCCL's version banner) to stderr
    global _JSON_OUT
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist

    import portablert_b200 as prt
    from portablert_b200 import hitreg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa_cpus = None
    if world > 1:
        from portablert_b200 import sharding
        if os.environ.get("PRT_BENCH_NUMA", "1") != "0":
            numa_cpus = sharding.bind_near_gpu(local)  # host slices are first-touched GPU-locally
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    backend = prt.CUDABackend(device=local)
    assert backend.is_available(), "bench.py needs a compute-capability-10.x GPU (no CPU fallback)"
    prt.select_backend(backend)

    # ---- inputs: rank 0 owns the scene; triangles are broadcast, each rank gets its ray slice
    def gpu_tracer(t, r):
        backend.set_tris(t)
        return backend.nearest_hits(r)

    # weak scaling: every rank traces its own copy of the SAME batch (per-GPU work is fixed and
    # identical, so the N-GPU value isolates system effects; distinct frames differ by up to 20 %
    # in cost because a handful of rays through the mesh's polar fans dominate the kernel tail)
    tris, rays, mask, desc = workload(args.config, frame=0, tracer=gpu_tracer)
    if world > 1:
        d_tris = torch.from_numpy(tris).to(dev) if rank == 0 else torch.empty(tris.shape, device=dev)
        dist.broadcast(d_tris, src=0)  # 36*N bytes over NVLink
    else:
        d_tris = torch.from_numpy(tris).to(dev)
    d_rays = torch.from_numpy(rays).to(dev)
    n_rays, n_tris = len(rays), len(tris)
    uv = torch.empty(n_rays, 2, device=dev)
    t = torch.empty(n_rays, device=dev)
    pid = torch.empty(n_rays, dtype=torch.int32, device=dev)
    p = torch.empty(n_rays, 3, device=dev)
    valid = torch.empty(n_rays, dtype=torch.uint8, device=dev)
    outs = dict(uv=uv.data_ptr() if mask & 1 else 0, t=t.data_ptr() if mask & 2 else 0,
                pid=pid.data_ptr() if mask & 4 else 0, p=p.data_ptr() if mask & 8 else 0,
                valid=valid.data_ptr() if mask & 16 else 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def l2_flush():
        flush.zero_()
        torch.cuda.synchronize()

    # ---- build (set_tris on device-resident triangles), timed by the library's own CUDA events
    build_ms = []
    for k in range(args.warmup + args.steps):
        l2_flush()
        ms = backend.set_tris_dev(d_tris.data_ptr(), n_tris)
        if k >= args.warmup:
            build_ms.append(ms)
    # the same build with the tree optimisation inside set_tris (mode 1): what a static scene pays in
    # total, measured with all scratch allocated (the lazy default runs the same kernels later)
    tl_mode = int(os.environ.get("PRT_B200_TREELET_MODE", "3"))
    tl_passes = int(os.environ.get("PRT_B200_TREELET_PASSES", "2"))
    backend.set_tree_optimisation(1, tl_passes)
    build_opt_ms = []
    for k in range(3 + min(args.steps, 10)):
        l2_flush()
        ms = backend.set_tris_dev(d_tris.data_ptr(), n_tris)
        if k >= 3:
            build_opt_ms.append(ms)
    backend.set_tree_optimisation(tl_mode, tl_passes)
    backend.set_tris_dev(d_tris.data_ptr(), n_tris)

    # ---- traversal: W warm-up + exactly K timed steps
    # Static scenes (all configs but C5) are traced repeatedly: the library optimises their tree
    # lazily (treelet restructuring) once they have served max(32 rays per triangle, 8 Mi rays).
    # C5 is the dynamic scene: every step calls set_tris with the NEXT of four distinct frames of
    # the deforming height field and traces it once; with the default temporal reuse the library,
    # after the same lazy optimisation, refits the optimised topology in set_tris instead of
    # rebuilding.  The warm-up is extended until that steady state is reached; the one-off costs
    # and the per-frame set_tris time are reported under `build` (not part of the traversal time).
    dynamic = args.config == "c5"
    frames = [d_tris]
    if dynamic:
        from portablert_b200 import scenes as _sc
        frames += [torch.from_numpy(_sc.heightfield(f)).to(dev) for f in (1, 2, 3)]
    frame_no, set_tris_ms = [0], []

    def new_frame():
        if dynamic:
            frame_no[0] += 1
            set_tris_ms.append(backend.set_tris_dev(frames[frame_no[0] % len(frames)].data_ptr(), n_tris))

    for _ in range(args.warmup):
        new_frame()
        l2_flush()
        backend.trace_dev(d_rays.data_ptr(), n_rays, mask, **outs)
    extra_warmup, refits0 = 0, backend.refits

    def steady():
        if n_tris < 7 or tl_mode < 2:
            return True
        if dynamic:
            return tl_mode < 3 or backend.refits >= refits0 + 2
        return backend.tree_depth > 0

    while not steady() and extra_warmup < 64:
        new_frame()
        l2_flush()
        backend.trace_dev(d_rays.data_ptr(), n_rays, mask, **outs)
        extra_warmup += 1
    optimise_ms, tree_depth = backend.last_optimise_ms, backend.tree_depth
    set_tris_ms.clear()
    refits1, rejects1 = backend.refits, backend.refit_rejects
    sampler = ClockSampler(local).start() if rank == 0 else None
    barrier()
    launches0 = backend.launch_count
    wall0 = time.perf_counter()
    step_ms = []
    for _ in range(args.steps):
        new_frame()
        l2_flush()
        step_ms.append(backend.trace_dev(d_rays.data_ptr(), n_rays, mask, **outs))
    barrier()
    wall = time.perf_counter() - wall0
    launches = backend.launch_count - launches0  # (C5: includes the kernels of the per-frame set_tris)
    steady_set_tris_ms = float(np.mean(set_tris_ms)) if set_tris_ms else None
    refits_timed, rebuilds_timed = backend.refits - refits1, backend.refit_rejects - rejects1
    clocks = sampler.stop() if sampler else None
    dev_ms = float(np.sum(step_ms))
    if world > 1:
        tt = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms_max = float(tt.item())
        tb = torch.tensor([float(np.mean(build_ms))], device=dev, dtype=torch.float64)
        dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        build_ms_mean = float(tb.item())
    else:
        dev_ms_max = dev_ms
        build_ms_mean = float(np.mean(build_ms))
    ms_per_step = dev_ms_max / args.steps
    total_rays = n_rays * world
    value = total_rays / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: host buffers in, host HitReg records out (copies inside the timed region)
    stride = hitreg.layout(mask)[0]
    e2e_steps = max(3, min(args.steps, 10))
    if world == 1:
        # the reference-facing call itself: prt_b200_nearest_hits(host rays) -> host AoS records
        # (a) inputs in pinned host memory, result array reused: what the contract asks for
        from portablert_b200.backend import pinned_empty
        p_rays = pinned_empty(rays.shape, np.float32)
        p_rays[...] = rays
        p_hits = pinned_empty((n_rays,), hitreg.dtype(mask))
        for _ in range(2):
            backend.nearest_hits(p_rays, mask, out=p_hits)
        barrier()
        e2e_t0 = time.perf_counter()
        for _ in range(e2e_steps):
            new_frame()  # (C5: the frame's rebuild is part of its end-to-end time)
            hits = backend.nearest_hits(p_rays, mask, out=p_hits)
        barrier()
        e2e_s = (time.perf_counter() - e2e_t0) / e2e_steps
        # (b) pageable numpy in, fresh array out (what std::vector callers of the C++ API get)
        for _ in range(2):
            backend.nearest_hits(rays, mask)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            new_frame()
            hits = backend.nearest_hits(rays, mask)
        e2e_pageable_s = (time.perf_counter() - t0) / e2e_steps
        api = "prt_b200_nearest_hits (pinned host rays -> pinned host HitReg AoS; H2D, traversal and D2H pipelined over chunks)"
        # the PCIe floor of that call on this box: both directions run concurrently, so the call
        # cannot beat max(H2D bytes / H2D bandwidth, D2H bytes / D2H bandwidth)
        raw_in = torch.from_numpy(p_rays.view(np.uint8).reshape(-1))
        raw_out = torch.from_numpy(p_hits.view(np.uint8).reshape(-1))
        d_in = torch.empty(raw_in.numel(), dtype=torch.uint8, device=dev)
        d_out = torch.empty(raw_out.numel(), dtype=torch.uint8, device=dev)
        pcie = {}
        for name, dst, src in (("h2d", d_in, raw_in), ("d2h", raw_out, d_out)):
            best = float("inf")
            for _ in range(5):
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                dst.copy_(src, non_blocking=True)
                ev1.record()
                torch.cuda.synchronize()
                best = min(best, ev0.elapsed_time(ev1))
            pcie[name + "_gbs"] = src.numel() / best / 1e6
            pcie[name + "_ms"] = best
        pcie["floor_ms"] = max(pcie["h2d_ms"], pcie["d2h_ms"])
    else:
        # The whole N-frame batch lives in host shared memory; every rank page-locks its own
        # contiguous slice (first-touched on its GPU's NUMA node) and calls the host entry point
        # on it, so all PCIe links run in parallel and the hits land in ray order in the shared
        # result buffer.  The synthetic batch is N copies of one frame: each rank writes its copy.
        tag = "prt_b200_%s" % os.environ.get("MASTER_PORT", "0")
        if rank == 0:
            shm = sharding.SharedHostBatch(tag, total_rays, hitreg.dtype(mask), rank, world, True,
                                           first_touch=True)
        barrier()
        if rank != 0:
            shm = sharding.SharedHostBatch(tag, total_rays, hitreg.dtype(mask), rank, world, False,
                                           first_touch=True)
        shm.my_rays[:] = rays
        barrier()

        def e2e_step():
            backend.nearest_hits(shm.my_rays, mask, out=shm.my_hits)

        for _ in range(2):
            e2e_step()
        barrier()
        e2e_t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        e2e_s = (time.perf_counter() - e2e_t0) / e2e_steps
        te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te.item())
        # What the platform allows: every rank moves exactly its slice's bytes host->device and
        # device->host at the same time (two streams, no kernel), all ranks concurrently; the e2e
        # call cannot beat this.  On a virtualised 8-GPU box the PCIe/IOMMU path is shared, so this
        # floor -- not the GPUs -- decides how e2e scales with N.
        import ctypes as C
        cudart = C.CDLL("libcudart.so.12")
        raw_in = np.ascontiguousarray(shm.my_rays).view(np.uint8).reshape(-1) \
            if not shm.my_rays.flags["C_CONTIGUOUS"] else shm.my_rays.view(np.uint8).reshape(-1)
        raw_out = shm.my_hits.view(np.uint8).reshape(-1)
        d_in = torch.empty(raw_in.size, dtype=torch.uint8, device=dev)
        d_out = torch.empty(raw_out.size, dtype=torch.uint8, device=dev)
        s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        floor_s = float("inf")
        for _ in range(4):
            barrier()
            t0 = time.perf_counter()
            cudart.cudaMemcpyAsync(C.c_void_p(d_in.data_ptr()), C.c_void_p(raw_in.ctypes.data),
                                   C.c_size_t(raw_in.size), C.c_int(1), C.c_void_p(s_in.cuda_stream))
            cudart.cudaMemcpyAsync(C.c_void_p(raw_out.ctypes.data), C.c_void_p(d_out.data_ptr()),
                                   C.c_size_t(raw_out.size), C.c_int(2), C.c_void_p(s_out.cuda_stream))
            barrier()
            floor_s = min(floor_s, time.perf_counter() - t0)
        tf = torch.tensor([floor_s], device=dev, dtype=torch.float64)
        dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        pcie_concurrent_floor_ms = float(tf.item()) * 1e3
        if rank == 0:  # the gathered result really is the whole batch, in ray order
            e2e_valid_fraction = float(np.asarray(shm.hits["valid"]).mean()) if mask & 16 else None
        barrier()
        shm.close(unlink=(rank == 0))
        api = ("host batch in shared memory; each rank: prt_b200_nearest_hits on its page-locked "
               "contiguous slice (H2D, kernel, D2H) -> hits in ray order in the shared result")
    e2e = {"value": total_rays / e2e_s / 1e6, "unit": "Mrays/s",
           "h2d_bytes_per_step": 24 * total_rays, "d2h_bytes_per_step": stride * total_rays,
           "ms_per_step": e2e_s * 1e3, "api": api}
    if world > 1:
        e2e["pcie_concurrent_floor_ms"] = pcie_concurrent_floor_ms
        e2e["frac_of_pcie_floor"] = pcie_concurrent_floor_ms / (e2e_s * 1e3)
        e2e["pcie_note"] = ("floor = all ranks copying their slices H2D and D2H concurrently, no kernels "
                            "(best of 4, max over ranks, wall clock incl. one barrier)")
    if world == 1:
        e2e["pcie"] = pcie
        e2e["frac_of_pcie_floor"] = pcie["floor_ms"] / (e2e_s * 1e3)
        e2e["pageable_value"] = total_rays / e2e_pageable_s / 1e6
        e2e["pageable_note"] = ("same call with pageable numpy input and a freshly allocated "
                                "result (staged through pinned buffers by threaded memcpy)")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the traversal kernel (rank 0's slice)
    cnt = torch.zeros(n_rays, 2, dtype=torch.int32, device=dev)
    new_frame()
    torch.cuda.synchronize()
    backend.trace_count_dev(d_rays.data_ptr(), n_rays, cnt.data_ptr())
    c = cnt.to(torch.float64).mean(0).cpu().numpy()
    nodes_per_ray, tris_per_ray = float(c[0]), float(c[1])
    bytes_per_ray = 24 + 64 * nodes_per_ray + 64 * tris_per_ray + out_bytes(mask)
    kern_ms = float(np.mean(step_ms))
    achieved = n_rays * bytes_per_ray / (kern_ms * 1e-3) / 1e9
    peak, peak_src = measured_peaks()
    l2_gbs = backend.read_bandwidth(32 << 20, 50)
    hbm_read_gbs = backend.read_bandwidth(2 << 30, 3)
    compulsory = n_rays * (24 + out_bytes(mask)) + backend.bvh_bytes
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(args.config)
        except Exception:
            traffic = None
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "peak_source": peak_src, "kernel": "prt::k_trace<mask=%d,SoA>" % mask,
        "kernel_ms": kern_ms, "bytes_per_ray": bytes_per_ray, "nodes_per_ray": nodes_per_ray,
        "tris_per_ray": tris_per_ray, "l2_read_gbs_measured": l2_gbs,
        "hbm_read_gbs_measured": hbm_read_gbs, "frac_of_l2": achieved / l2_gbs,
        "compulsory_bytes": compulsory,
        "compulsory_frac_of_hbm": compulsory / (kern_ms * 1e-3) / 1e9 / peak,
        "note": "BVH (%.1f MB) is L2-resident: fetched bytes are served by L1/L2, so the HBM "
                "fraction can exceed the DRAM traffic; frac_of_l2 is the binding roofline" %
                (backend.bvh_bytes / 1e6),
    }
    # algorithmic HBM bytes per triangle: bounds 36 R; morton 36 R + 12 W; histograms
    # 8 R; P sort passes x (12 R + 12 W); hierarchy kernel: 4 (index) + 36 (triangle) + 16 (keys) R,
    # 64 W (record), 64 W (node halves), 24 R (sibling box), 8 (bound exchange)
    passes = 4 if n_tris <= (1 << 16) else (5 if n_tris <= (1 << 22) else 6)
    build_bytes_per_tri = 36 + 48 + 8 + 24 * passes + 56 + 64 + 64 + 24 + 8
    build = {"mtris_s": n_tris / (build_ms_mean * 1e-3) / 1e6, "ms": build_ms_mean,
             "ms_with_optimisation": float(np.mean(build_opt_ms)),
             "mtris_s_with_optimisation": n_tris / (float(np.mean(build_opt_ms)) * 1e-3) / 1e6,
             "optimise_passes": tl_passes,
             "first_lazy_optimise_ms": optimise_ms, "tree_height": tree_depth,
             "set_tris_ms_steady": steady_set_tris_ms, "refits_in_timed_steps": refits_timed,
             "rebuilds_in_timed_steps": rebuilds_timed,
             "tree": (("deforming mesh, 4 distinct frames in rotation: after the lazy optimisation set_tris "
                       "refits the optimised topology (temporal reuse, mode 3; `set_tris_ms_steady`); `ms` is "
                       "a full rebuild of the plain LBVH" if refits_timed else
                       "plain LBVH, rebuilt before every step (dynamic scene)") if dynamic else
                      "LBVH from set_tris (timed as `ms`), then optimised once by treelet "
                      "restructuring after max(32 rays per triangle, 8 Mi rays) (inside the warm-up, "
                      "which was extended by %d steps for it; `first_lazy_optimise_ms` includes "
                      "first-use allocations, `ms_with_optimisation` is the steady-state cost of "
                      "build + optimisation)" % extra_warmup
                      if optimise_ms > 0 else "plain LBVH"),

             "bytes_per_tri": build_bytes_per_tri,
             "achieved_gbs": n_tris * build_bytes_per_tri / (build_ms_mean * 1e-3) / 1e9,
             "frac_of_hbm": n_tris * build_bytes_per_tri / (build_ms_mean * 1e-3) / 1e9 / peak}

    per_mask = None
    if args.per_mask:
        per_mask = {}
        for combo in hitreg.TAG_COMBOS:
            m = hitreg.mask_of(combo)
            o = dict(uv=uv.data_ptr() if m & 1 else 0, t=t.data_ptr() if m & 2 else 0,
                     pid=pid.data_ptr() if m & 4 else 0, p=p.data_ptr() if m & 8 else 0,
                     valid=valid.data_ptr() if m & 16 else 0)
            ts = []
            for k in range(6):
                l2_flush()
                ms = backend.trace_dev(d_rays.data_ptr(), n_rays, m, **o)
                if k >= 2:
                    ts.append(ms)
            per_mask["_".join(combo)] = n_rays / (np.mean(ts) * 1e-3) / 1e6

    # ---- opt-in watertight mode beside the default (reference arithmetic): its throughput and how
    # many rays of this batch it answers differently
    watertight = None
    if world == 1 and mask & 2:
        if dynamic:  # compare on one and the same frame
            backend.set_tris_dev(d_tris.data_ptr(), n_tris)
        backend.trace_dev(d_rays.data_ptr(), n_rays, mask, **outs)
        t_def = t.clone()
        backend.set_triangle_test(1)
        # same tree state as the timed steps: optimised for static scenes, plain for the dynamic one
        backend.set_tree_optimisation(0 if tree_depth == 0 else 1, 2)
        backend.set_tris_dev(d_tris.data_ptr(), n_tris)
        ts = []
        for k in range(7):
            l2_flush()
            ms = backend.trace_dev(d_rays.data_ptr(), n_rays, mask, **outs)
            if k >= 2:
                ts.append(ms)
        hit_def, hit_wt = torch.isfinite(t_def), torch.isfinite(t)
        both = hit_def & hit_wt
        rel = ((t[both] - t_def[both]).abs() / t_def[both].abs().clamp_min(1e-30))
        watertight = {"value": n_rays / (np.mean(ts) * 1e-3) / 1e6, "unit": "Mrays/s",
                      "valid_differs": int((hit_def != hit_wt).sum().item()),
                      "t_differs": int((t[both] != t_def[both]).sum().item()),
                      "t_maxrel": float(rel.max().item()) if rel.numel() else 0.0,
                      "t_rel_gt_1e-5": int((rel > 1e-5).sum().item()),
                      "note": "PRT_B200_WATERTIGHT=1 (Woop et al. 2013) vs the default on the same rays"}
        backend.set_triangle_test(0)
        backend.set_tree_optimisation(tl_mode, tl_passes)
        backend.set_tris_dev(d_tris.data_ptr(), n_tris)

    # ---- C5 without temporal reuse (mode 2), same frames: every set_tris rebuilds the plain LBVH
    dynamic_plain = None
    if world == 1 and dynamic:
        backend.set_tree_optimisation(2, tl_passes)
        b_ms, t_ms = [], []
        for i in range(2 + max(args.steps, 8)):
            l2_flush()
            b = backend.set_tris_dev(frames[i % len(frames)].data_ptr(), n_tris)
            l2_flush()
            t_ = backend.trace_dev(d_rays.data_ptr(), n_rays, mask, **outs)
            if i >= 2:
                b_ms.append(b)
                t_ms.append(t_)
        dynamic_plain = {
            "set_tris_ms": float(np.mean(b_ms)), "trace_ms": float(np.mean(t_ms)),
            "frame_ms": float(np.mean(b_ms) + np.mean(t_ms)),
            "value": n_rays / (float(np.mean(t_ms)) * 1e-3) / 1e6, "unit": "Mrays/s",
            "default_frame_ms": (steady_set_tris_ms or 0.0) + ms_per_step,
            "note": "PRT_B200_TREELET_MODE=2 (no temporal reuse): the same 4 frames, plain LBVH rebuilt "
                    "by every set_tris; default_frame_ms = set_tris_ms_steady + ms_per_step of the main run"}
        backend.set_tree_optimisation(tl_mode, tl_passes)
        backend.set_tris_dev(d_tris.data_ptr(), n_tris)

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_reference(tris, rays, mask, budget_s=20.0)
        except Exception as e:  # the checker is optional equipment for the bench
            cpu = {"unavailable": str(e)}

    line = {
        "metric": "nearest_hits throughput", "value": value, "unit": "Mrays/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": desc, "rays_per_gpu": n_rays, "tris": n_tris, "tag_mask": mask,
                   "l2": "flushed between timed steps (256 MiB memset)",
                   "multi_gpu": "tris NCCL-broadcast from rank 0, identical BVH built on every rank, "
                                "batch = N copies of the workload's rays, rank r traces the r-th contiguous slice; e2e: "
                                "batch in host shared memory, each rank DMA's its own slice, hits "
                                "land in ray order in the shared result",
                   "numa": (None if world == 1 else
                            "rank pinned to %d GPU-local CPUs, slices first-touched there"
                            % len(numa_cpus) if numa_cpus else "no NUMA binding")},
        "build_mtris_s": build["mtris_s"], "build": build,
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "roofline": roofline, "cpu_baseline": cpu, "watertight": watertight, "dynamic_without_reuse": dynamic_plain,
        "wall_s_timed_region": wall, "device": backend.device_name(),
    }
    if per_mask:
        line["per_mask_mrays_s"] = per_mask
    print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
