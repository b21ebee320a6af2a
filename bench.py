#!/usr/bin/env python
"""bench.py -- device-timed Mrays/s of nearest_hits (+ BVH build Mtris/s) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c4|c2|c3|c3b|c5] [--impl reference]

One JSON line on stdout (rank 0).  A "step" is one nearest_hits pass over the whole ray batch of
the workload.  Default workload = BASELINE.json configs[3] (C4), the configuration the metric is
quoted on: procedural 10 M-triangle scene, 100 M incoherent rays, tags t+primitive_id, STRONG-scaled
over N GPUs (rank r traces the r-th contiguous slice of the same 100 M rays).

  value       whole-job Mrays/s, rays resident in HBM, SoA outputs in HBM; per-step device time
              from CUDA events on the stream the kernels run on (taken by the library around ray
              reordering + traversal), L2 flushed between steps, max over ranks
  e2e         the same metric through the reference-facing host API (host ray buffer in, host
              HitReg records out: H2D + kernels + D2H inside the timed region); at N > 1 it is ONE
              process driving all N GPUs through a multi-GPU context of the library
              (prt_b200_create_multi: triangles broadcast over NVLink, the host batch cut into
              contiguous slices, hits written in ray order) -- what a C++ user of the reference
              API gets with PRT_B200_GPUS=N.  e2e.cxx_plugin = the reference's real C++ signature
              (std::vector in, fresh std::vector out) timed by oracle/_ref/bench_cxx.
  roofline    the traversal kernel against the roof that binds it -- measured HBM copy bandwidth
              when the BVH exceeds L2, the L2 read bandwidth measured in this run otherwise --
              from ALGORITHMIC bytes per ray counted by the instrumented kernel:
              24 (ray) + 64 * nodes fetched + 64 * triangles tested + output bytes
  parity      the section-8c comparator against the unmodified reference CPU backend on a fixed
              sample of the batch (prefix + strided), run inside this bench (N = 1)
  cpu_baseline  that reference (oracle/_ref) timed on the box's host cores
  --impl reference  times only that CPU reference and prints the same line shape

Other configs: --config c2 (69 k-tri mesh, 1920x1080 primary, all tags), c3 (262 k-tri interior,
4K primary), c3b (its one-bounce diffuse rays), c5 (1 M-tri height field rebuilt every frame, 8 M
rays).  PRT_BENCH_C4_SPHERES / PRT_BENCH_C4_RAYS scale C4 down for experiments.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# ------------------------------------------------------------------------------------------------
def workload(name: str, frame: int = 0, tracer=None, part=(0, 1)):
    """-> dict(tris (N,9) f32, rays: this part's contiguous slice (R,6) f32, lo, n_total, mask,
    desc).  part = (rank, world): ray i of n belongs to part floor(i * world / n)."""
    from portablert_b200 import hitreg, scenes
    rank, world = part

    def cut(rays):
        n = len(rays)
        lo, hi = n * rank // world, n * (rank + 1) // world
        return np.ascontiguousarray(rays[lo:hi]), lo, n

    if name == "c2":
        tris = scenes.blob()
        rays, lo, n = cut(scenes.pinhole_rays(1920, 1080, cam=(0.002 * frame, 0.0, -0.3)))
        mask, desc = hitreg.ALL, ("C2: 69192-tri displaced-sphere mesh, 1920x1080 pinhole "
                                  "primary rays, tags uv,t,primitive_id,p,valid")
    elif name == "c3":
        tris = scenes.interior()
        rays, lo, n = cut(scenes.camera_rays(3840, 2160, (2 + 0.01 * frame, 6, 3), (28, 4, 15)))
        mask, desc = hitreg.ALL, "C3: %d-tri interior, 3840x2160 primary rays, all tags" % len(tris)
    elif name == "c3b":
        # one-bounce incoherent diffuse rays spawned at the primary hits of C3 (the primary hits
        # are input preparation: `tracer(tris, rays) -> full hit records` runs before any timing)
        tris = scenes.interior()
        prim = scenes.camera_rays(3840, 2160, (2 + 0.01 * frame, 6, 3), (28, 4, 15))
        h = tracer(tris, prim)
        p = np.stack([h["px"], h["py"], h["pz"]], -1)
        rays, lo, n = cut(scenes.bounce_rays(tris, prim, h["valid"], h["primitive_id"], p)[0])
        mask, desc = hitreg.T | hitreg.PID, ("C3 bounce: %d-tri interior, %d one-bounce cosine-"
                                             "weighted diffuse rays, t+primitive_id" % (len(tris), n))
    elif name == "c5":
        tris = scenes.heightfield(frame)
        rays, lo, n = cut(scenes.camera_rays(3840, 2160, (10, 6, -4), (10, 0, 5))[:8_000_000])
        mask, desc = hitreg.T | hitreg.VALID, "C5: 1M-tri heightfield, 8M primary rays, t+valid"
    elif name == "c4":
        n_s = int(os.environ.get("PRT_BENCH_C4_SPHERES", "10000"))
        n = int(os.environ.get("PRT_BENCH_C4_RAYS", "100000000"))
        tris = scenes.sphere_field(n_s)
        blo, bhi = tris.reshape(-1, 3).min(0), tris.reshape(-1, 3).max(0)
        lo, hi = n * rank // world, n * (rank + 1) // world
        rays = scenes.incoherent_rays_range(lo, hi - lo, blo, bhi, seed=4 + frame)
        mask, desc = hitreg.T | hitreg.PID, ("C4: %d-tri sphere field, %d incoherent rays, "
                                             "t+primitive_id" % (len(tris), n))
    else:
        raise SystemExit(f"unknown config {name}")
    return dict(tris=tris, rays=rays, lo=lo, n_total=n, mask=mask, desc=desc)


def config_dict(w, args):
    """The part of `config` both arms print (the driver compares the key sets)."""
    return {"workload": w["desc"], "rays": int(w["n_total"]), "tris": int(len(w["tris"])),
            "tag_mask": int(w["mask"]), "name": args.config,
            "l2": "flushed between timed steps (256 MiB memset); inputs larger than L2",
            "multi_gpu": "strong scaling: rank r traces the contiguous slice [r*R/N, (r+1)*R/N) of "
                         "the same R rays; triangles NCCL-broadcast from rank 0, identical BVH "
                         "built on every rank, no collective on the data path"}


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML DURING the timed region (a thread
    polling every ~1 ms; nvidia-smi's 100 ms floor cannot see a millisecond-scale region)."""

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
            time.sleep(0.001)

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


def sample_indices(n, per=1_000_000):
    """The fixed parity sample of a batch: its first `per` rays plus a strided `per` over all of it."""
    if n <= 2 * per:
        return np.arange(n, dtype=np.int64)
    return np.unique(np.concatenate([np.arange(per, dtype=np.int64),
                                     np.arange(0, n, n // per, dtype=np.int64)]))


def checksum(t=None, pid=None, valid=None):
    """Order-independent digest of a result: (#valid, sum of primitive ids of the hits, xor of the
    t bit patterns) -- the same three numbers oracle/_ref/bench_cxx prints."""
    import torch
    if valid is None:
        valid = torch.isfinite(t) if isinstance(t, torch.Tensor) else np.isfinite(t)
    if isinstance(valid, np.ndarray):
        v = valid.astype(bool)
        n_valid = int(v.sum())
        psum = int(pid[v].astype(np.uint64).sum()) if pid is not None else 0
        x = int(np.bitwise_xor.reduce(np.ascontiguousarray(t).view(np.uint32))) if t is not None and len(t) else 0
        return n_valid, psum, x
    v = valid.bool()
    n_valid = int(v.sum().item())
    psum = int(pid.to(torch.int64).masked_fill(~v, 0).bitwise_and(0xffffffff).sum().item()) if pid is not None else 0
    x = 0
    if t is not None and t.numel():
        b = t.view(torch.int32)
        while b.numel() > 1:  # xor-reduce (no torch primitive for it)
            if b.numel() & 1:
                b = torch.cat([b, b.new_zeros(1)])
            b = b[: b.numel() // 2] ^ b[b.numel() // 2:]
        x = int(b.item()) & 0xffffffff
    return n_valid, psum, x


# ------------------------------------------------------------------------------------------------
class CpuReference:
    """The unmodified reference CPU backend (oracle/_ref; test infrastructure, used here as the
    checker and the reported baseline only).  set_tris runs on a background thread (ctypes drops
    the GIL; the reference's build is single-threaded: 10 M triangles take about two minutes)
    while the GPU legs run."""

    def __init__(self, tris):
        from oracle import Reference
        self.ref = Reference()
        self.tris = tris
        self.err = None
        self.build_s = None
        self.thread = threading.Thread(target=self._build, daemon=True)
        self.thread.start()

    def _build(self):
        try:
            self.build_s = self.ref.set_tris(self.tris)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def wait(self):
        self.thread.join()
        if self.err:
            raise RuntimeError(self.err)

    def trace(self, rays, mask, keep=True):
        hits = self.ref.nearest_hits(rays, mask, keep=keep)
        return hits, self.ref.last_trace_s

    def timed(self, rays, mask, budget_s, steps=1, warmup=0):
        """Mrays/s on a bounded sample of `rays` (about budget_s of CPU work in total)."""
        ref = self.ref
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
        return {"value": len(sample) / s / 1e6, "unit": "Mrays/s", "cores": ref.threads,
                "kind": "reference",
                "sample": f"{what}; set_tris on all {len(self.tris)} tris (1 thread)",
                "build_mtris_s": len(self.tris) / self.build_s / 1e6, "build_s": self.build_s,
                "trace_s": s, "cpu": ref.device_name(), "n_sample": int(len(sample))}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return

    def ref_tracer(t, r):
        from oracle import Reference
        ref = Reference()
        ref.set_tris(t)
        return ref.nearest_hits(r, 31)

    w = workload(args.config, tracer=ref_tracer)
    cpu = CpuReference(w["tris"])
    cpu.wait()
    cb = cpu.timed(w["rays"], w["mask"], budget_s=60.0, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": "nearest_hits throughput", "value": cb["value"],
        "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": cb["trace_s"] * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(w, args),
        "build_mtris_s": cb["build_mtris_s"],
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)


# ------------------------------------------------------------------------------------------------
def cxx_plugin_leg(w, rays, steps):
    """e2e through the reference's real C++ signature: oracle/_ref/bench_cxx (the reference's
    headers + this backend in one binary) reads the workload from files and times
    nearest_hits<Tags...>(std::vector<Ray>) -> std::vector<HitReg<Tags...>>."""
    exe = os.path.join(ROOT, "oracle", "_ref", "bench_cxx")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/bench_cxx not built (needs the reference checkout)"}
    need = rays.nbytes + w["tris"].nbytes
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 0
    n = len(rays)
    if avail and avail < 6 * need:  # the program holds rays + a fresh result per call
        n = max(1, int(n * avail / (6.0 * need)))
    tmp = None
    for base in ("/dev/shm", tempfile.gettempdir()):
        try:
            if shutil.disk_usage(base).free > 1.2 * need:
                tmp = tempfile.mkdtemp(prefix="prt_b200_", dir=base)
                break
        except Exception:
            continue
    if tmp is None:
        return {"unavailable": "no room for the workload files"}
    try:
        tp, rp = os.path.join(tmp, "tris.bin"), os.path.join(tmp, "rays.bin")
        w["tris"].tofile(tp)
        rays[:n].tofile(rp)
        # warm-up: enough calls for the library's lazy tree optimisation (after max(32 rays per
        # triangle, 8 Mi rays)) and its first-use allocations to lie outside the timed calls
        warm = 2 + int(np.ceil(max(32 * len(w["tris"]), 8 << 20) / max(1, n)))
        out = subprocess.run([exe, tp, rp, str(w["mask"]), str(steps), str(min(warm, 12))],
                             capture_output=True, text=True, timeout=900)
        lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
        if out.returncode != 0 or not lines:
            return {"unavailable": "bench_cxx failed: " + (out.stderr or out.stdout)[-300:]}
        res = json.loads(lines[-1])
        res["n_rays"] = n
        return res
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def embree_status():
    """The reference's Embree CPU backend is the second oracle `north_star` names.  Embree 4 is an
    external binary dependency of the reference (README.md:71-76, CMakeLists.txt:103-115) that this
    image does not carry; tests/dropin/Makefile wires the backend into the acceptance program
    whenever <embree4/rtcore.h> is found, and this line says which of the two happened."""
    exe = os.path.join(ROOT, "oracle", "_ref", "acceptance")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/acceptance not built"}
    try:
        out = subprocess.run([exe, "--list"], capture_output=True, text=True, timeout=120).stdout
    except Exception as e:  # noqa: BLE001
        return {"unavailable": repr(e)}
    if "Embree CPU backend: compiled in" in out:
        return {"available": True, "note": "compared by oracle/_ref/acceptance (results.csv)"}
    return {"unavailable": "Embree 4 not installed when oracle/_ref/acceptance was built "
                           "(USE_EMBREE_CPU wiring: tests/dropin/Makefile)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="c4")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip watertight / C++ plugin legs")
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
    from portablert_b200.backend import pinned_empty

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    host_group = None
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        host_group = dist.new_group(backend="gloo")  # host-side barriers that leave the GPUs alone
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    backend = prt.CUDABackend(device=local)
    assert backend.is_available(), "bench.py needs a compute-capability-10.x GPU (no CPU fallback)"
    prt.select_backend(backend)

    def gpu_tracer(t, r):
        backend.set_tris(t)
        return backend.nearest_hits(r)

    # ---- inputs: strong scaling -- this rank's contiguous slice of the one batch
    t_gen = time.perf_counter()
    w = workload(args.config, frame=0, tracer=gpu_tracer, part=(rank, world))
    tris, rays, mask = w["tris"], w["rays"], w["mask"]
    gen_s = time.perf_counter() - t_gen
    n_total, n_rays, n_tris = w["n_total"], len(rays), len(tris)
    if world > 1:  # the scene reaches the other GPUs over NVLink, as it would from a loader rank
        d_tris = torch.from_numpy(tris).to(dev) if rank == 0 else torch.empty(tris.shape, device=dev)
        dist.broadcast(d_tris, src=0)
    else:
        d_tris = torch.from_numpy(tris).to(dev)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu = CpuReference(tris)  # builds in the background while the GPU legs run
        except Exception as e:  # the checker is optional equipment for the bench
            cpu = None
            cpu_err = str(e)
    d_rays = torch.from_numpy(rays).to(dev)
    uv = torch.empty((n_rays, 2) if mask & 1 else (1, 2), device=dev)
    t = torch.empty(n_rays if mask & 2 else 1, device=dev)
    pid = torch.empty(n_rays if mask & 4 else 1, dtype=torch.int32, device=dev)
    p = torch.empty((n_rays, 3) if mask & 8 else (1, 3), device=dev)
    valid = torch.empty(n_rays if mask & 16 else 1, dtype=torch.uint8, device=dev)
    outs = dict(uv=uv.data_ptr() if mask & 1 else 0, t=t.data_ptr() if mask & 2 else 0,
                pid=pid.data_ptr() if mask & 4 else 0, p=p.data_ptr() if mask & 8 else 0,
                valid=valid.data_ptr() if mask & 16 else 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def host_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=host_group)

    def l2_flush():
        flush.zero_()
        torch.cuda.synchronize()

    # ---- build (set_tris on device-resident triangles), timed by the library's own CUDA events
    build_ms = []
    n_build = min(args.steps, 10)
    for k in range(3 + n_build):
        l2_flush()
        ms = backend.set_tris_dev(d_tris.data_ptr(), n_tris)
        if k >= 3:
            build_ms.append(ms)
    # the same build with the tree optimisation inside set_tris (mode 1): what a static scene pays in
    # total, measured with all scratch allocated (the lazy default runs the same kernels later)
    tl_mode = int(os.environ.get("PRT_B200_TREELET_MODE", "3"))
    tl_passes = int(os.environ.get("PRT_B200_TREELET_PASSES", "2"))
    backend.set_tree_optimisation(1, tl_passes)
    build_opt_ms = []
    for k in range(2 + min(args.steps, 5)):
        l2_flush()
        ms = backend.set_tris_dev(d_tris.data_ptr(), n_tris)
        if k >= 2:
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

    # (strong scaling: the lazy optimisation counts rays per GPU, so the warm-up is extended until
    # every rank's tree is optimised -- the steady state of a static scene)
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
    step_ms, kern_ms_list = [], []
    for _ in range(args.steps):
        new_frame()
        l2_flush()
        step_ms.append(backend.trace_dev(d_rays.data_ptr(), n_rays, mask, **outs))
        kern_ms_list.append(backend.last_kernel_ms)
    barrier()
    wall = time.perf_counter() - wall0
    launches = backend.launch_count - launches0  # (C5: includes the kernels of the per-frame set_tris)
    steady_set_tris_ms = float(np.mean(set_tris_ms)) if set_tris_ms else None
    refits_timed, rebuilds_timed = backend.refits - refits1, backend.refit_rejects - rejects1
    clocks = sampler.stop() if sampler else None
    dev_ms = float(np.sum(step_ms))
    if world > 1:
        tt = torch.tensor([dev_ms, float(np.mean(build_ms)), float(launches)], device=dev, dtype=torch.float64)
        tmax = tt.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        dev_ms_max, build_ms_mean, launches = float(tmax[0]), float(tmax[1]), int(tt[2].item())
    else:
        dev_ms_max = dev_ms
        build_ms_mean = float(np.mean(build_ms))
    ms_per_step = dev_ms_max / args.steps
    value = n_total / (ms_per_step * 1e-3) / 1e6

    # ---- device results of this rank's slice: digest for the cross-checks below
    if dynamic:
        backend.set_tris_dev(d_tris.data_ptr(), n_tris)
        backend.trace_dev(d_rays.data_ptr(), n_rays, mask, **outs)
    dev_sum = checksum(t if mask & 2 else None, pid if mask & 4 else None,
                       valid if (mask & 16) and not (mask & 2) else None)
    replicas = None
    if world > 1:
        # (1) every rank traces the SAME leading rays (rank 0's first million) on its own replica
        # of the BVH: the digests must agree bit for bit across GPUs
        ns = min(1_000_000, n_total // world)
        lead = workload(args.config, part=(0, world))["rays"][:ns] if args.config == "c4" else None
        if lead is not None:
            dl = torch.from_numpy(lead).to(dev)
            t2 = torch.empty(ns, device=dev)
            p2 = torch.empty(ns, dtype=torch.int32, device=dev)
            backend.trace_dev(dl.data_ptr(), ns, 6, t=t2.data_ptr(), pid=p2.data_ptr())
            mine = torch.tensor(checksum(t2, p2), dtype=torch.int64, device=dev)
            allv = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allv, mine)
            replicas = {"rays": ns, "digests_identical_across_gpus":
                        bool(all(torch.equal(a, allv[0]) for a in allv))}
            del dl, t2, p2
        # (2) the digest of the whole batch = combination of the slices' digests
        parts = torch.tensor(dev_sum, dtype=torch.int64, device=dev)
        gathered = [torch.zeros_like(parts) for _ in range(world)]
        dist.all_gather(gathered, parts)
        whole = [int(sum(int(g[0]) for g in gathered)), int(sum(int(g[1]) for g in gathered)), 0]
        for g in gathered:
            whole[2] ^= int(g[2])
        dev_sum = tuple(whole)

    # ---- e2e: host buffers in, host HitReg records out (copies inside the timed region)
    stride = hitreg.layout(mask)[0]
    e2e_steps = max(3, min(args.steps, 10))
    hit_dt = hitreg.dtype(mask)

    def digest_hits(h):
        return checksum(h["t"] if mask & 2 else None, h["primitive_id"] if mask & 4 else None,
                        h["valid"] if (mask & 16) and not (mask & 2) else None)

    e2e = None
    if world == 1:
        # the reference-facing call itself: prt_b200_nearest_hits(host rays) -> host AoS records
        # (a) inputs in pinned host memory, result array reused: what the contract asks for
        p_rays = pinned_empty(rays.shape, np.float32)
        p_rays[...] = rays
        p_hits = pinned_empty((n_rays,), hit_dt)
        for _ in range(2):
            backend.nearest_hits(p_rays, mask, out=p_hits)
        barrier()
        e2e_t0 = time.perf_counter()
        for _ in range(e2e_steps):
            new_frame()  # (C5: the frame's rebuild is part of its end-to-end time)
            backend.nearest_hits(p_rays, mask, out=p_hits)
        barrier()
        e2e_s = (time.perf_counter() - e2e_t0) / e2e_steps
        h2d_b, d2h_b = backend.last_transfer_bytes
        if dynamic:
            backend.set_tris_dev(d_tris.data_ptr(), n_tris)
            backend.nearest_hits(p_rays, mask, out=p_hits)
        e2e_digest_ok = digest_hits(p_hits) == dev_sum
        # (b) pageable numpy in, fresh array out (what std::vector callers of the C++ API get)
        for _ in range(2):
            backend.nearest_hits(rays, mask)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            new_frame()
            hits = backend.nearest_hits(rays, mask)
        e2e_pageable_s = (time.perf_counter() - t0) / e2e_steps
        pg_h2d, pg_d2h = backend.last_transfer_bytes
        del hits
        api = ("prt_b200_nearest_hits (pinned host rays -> pinned host HitReg AoS; H2D, traversal "
               "and D2H pipelined over chunks)")
        # the PCIe floor of that call on this box: both directions run concurrently, so the call
        # cannot beat max(H2D bytes / H2D bandwidth, D2H bytes / D2H bandwidth)
        raw_in = torch.from_numpy(p_rays.view(np.uint8).reshape(-1))
        raw_out = torch.from_numpy(p_hits.view(np.uint8).reshape(-1))
        d_in = torch.empty(raw_in.numel(), dtype=torch.uint8, device=dev)
        d_out = torch.empty(raw_out.numel(), dtype=torch.uint8, device=dev)
        pcie = {}
        for name, dst, src in (("h2d", d_in, raw_in), ("d2h", raw_out, d_out)):
            best = float("inf")
            for _ in range(3):
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                dst.copy_(src, non_blocking=True)
                ev1.record()
                torch.cuda.synchronize()
                best = min(best, ev0.elapsed_time(ev1))
            pcie[name + "_gbs"] = src.numel() / best / 1e6
            pcie[name + "_ms"] = best
        pcie["floor_ms"] = max(pcie["h2d_ms"], pcie["d2h_ms"])
        del d_in, d_out, raw_in, raw_out
        e2e = {"value": n_total / e2e_s / 1e6, "unit": "Mrays/s",
               "h2d_bytes_per_step": int(h2d_b), "d2h_bytes_per_step": int(d2h_b),
               "ms_per_step": e2e_s * 1e3, "api": api, "digest_equals_device_result": e2e_digest_ok,
               "pcie": pcie, "frac_of_pcie_floor": pcie["floor_ms"] / (e2e_s * 1e3),
               "pageable_value": n_total / e2e_pageable_s / 1e6,
               "pageable_ms_per_step": e2e_pageable_s * 1e3,
               "pageable_h2d_bytes": int(pg_h2d), "pageable_d2h_bytes": int(pg_d2h),
               "pageable_note": ("same call with pageable numpy input and a freshly allocated "
                                 "result: staged through pinned buffers by pooled copy threads, "
                                 "results cross PCIe tightly packed and are scattered into the "
                                 "records on the way out")}
        if not args.no_extras and not dynamic:
            cx = cxx_plugin_leg(w, rays, max(3, min(args.steps, 5)))
            if "ms_mean" in cx:
                cx["value"] = cx["n_rays"] / (cx["ms_mean"] * 1e-3) / 1e6
                if cx["n_rays"] == n_total:
                    cx["digest_equals_device_result"] = \
                        (cx["valid"], cx["pid_sum"], cx["t_xor"]) == dev_sum
            e2e["cxx_plugin"] = cx
        del p_rays, p_hits
    else:
        # ONE process (rank 0) drives all N GPUs through a multi-GPU context of the library: the
        # whole batch sits in its pinned host memory, the library broadcasts the triangles, cuts
        # the batch into N contiguous slices, runs one copy pipeline per GPU and writes the hits
        # in ray order.  The other ranks have released nothing (their replicas stay resident) and
        # wait at a host-side barrier.
        host_barrier()
        if rank == 0:
            try:
                os.sched_setaffinity(0, range(os.cpu_count()))
            except Exception:
                pass
            full = workload(args.config, part=(0, 1))["rays"] if args.config != "c3b" else None
            if full is None:
                e2e = {"unavailable": "c3b needs a tracer to make its rays"}
            else:
                mg = prt.CUDABackend(gpus=world)
                mg.init()
                mg.set_tris(tris)
                p_rays = pinned_empty(full.shape, np.float32)
                p_rays[...] = full
                del full
                p_hits = pinned_empty((n_total,), hit_dt)
                for _ in range(2 + (2 if tl_mode >= 2 else 0)):
                    mg.nearest_hits(p_rays, mask, out=p_hits)
                # (a static scene in steady state: optimise now rather than wait for the lazy rule)
                t0 = time.perf_counter()
                for _ in range(e2e_steps):
                    mg.nearest_hits(p_rays, mask, out=p_hits)
                e2e_s = (time.perf_counter() - t0) / e2e_steps
                h2d_b, d2h_b = mg.last_transfer_bytes
                ok = digest_hits(p_hits) == dev_sum
                # what the platform allows: all slices copied H2D and D2H at the same time on all
                # GPUs, no kernels (the shared host PCIe / IOMMU path decides how e2e scales with N)
                bounds = [(n_total * i // world, n_total * (i + 1) // world) for i in range(world)]
                raw_in = torch.from_numpy(p_rays.view(np.uint8).reshape(-1))
                raw_out = torch.from_numpy(p_hits.view(np.uint8).reshape(-1))
                bufs = []
                for i, (a, b) in enumerate(bounds):
                    dv = torch.device("cuda", i)
                    bufs.append((torch.empty((b - a) * 24, dtype=torch.uint8, device=dv),
                                 torch.empty((b - a) * stride, dtype=torch.uint8, device=dv),
                                 torch.cuda.Stream(dv), torch.cuda.Stream(dv)))
                floor_s = float("inf")
                for _ in range(3):
                    for i in range(world):
                        torch.cuda.synchronize(i)
                    t0 = time.perf_counter()
                    for i, (a, b) in enumerate(bounds):
                        din, dout, s1, s2 = bufs[i]
                        with torch.cuda.stream(s1):
                            din.copy_(raw_in[a * 24:b * 24], non_blocking=True)
                        with torch.cuda.stream(s2):
                            raw_out[a * stride:b * stride].copy_(dout, non_blocking=True)
                    for i in range(world):
                        torch.cuda.synchronize(i)
                    floor_s = min(floor_s, time.perf_counter() - t0)
                torch.cuda.set_device(dev)
                e2e = {"value": n_total / e2e_s / 1e6, "unit": "Mrays/s",
                       "h2d_bytes_per_step": int(h2d_b), "d2h_bytes_per_step": int(d2h_b),
                       "ms_per_step": e2e_s * 1e3,
                       "api": ("ONE process, multi-GPU context of the library (prt_b200_create_multi, "
                               "%d GPUs): prt_b200_nearest_hits on the whole pinned host batch -> "
                               "contiguous slice per GPU, hits in ray order in the caller's array; "
                               "triangles broadcast device 0 -> others via %s"
                               % (mg.num_devices, mg.broadcast_path or "-")),
                       "digest_equals_per_rank_device_results": ok,
                       "pcie_concurrent_floor_ms": floor_s * 1e3,
                       "frac_of_pcie_floor": floor_s / e2e_s,
                       "pcie_note": ("floor = all GPUs copying their slices H2D and D2H concurrently "
                                     "from this process, no kernels (best of 3, wall clock)")}
                del bufs, raw_in, raw_out, p_rays, p_hits
                mg.shutdown()
        host_barrier()

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
    del cnt
    nodes_per_ray, tris_per_ray = float(c[0]), float(c[1])
    bytes_per_ray = 24 + 64 * nodes_per_ray + 64 * tris_per_ray + out_bytes(mask)
    kern_ms = float(np.mean(kern_ms_list))
    achieved = n_rays * bytes_per_ray / (kern_ms * 1e-3) / 1e9
    hbm_peak, peak_src = measured_peaks()
    l2_gbs = backend.read_bandwidth(32 << 20, 50)
    hbm_read_gbs = backend.read_bandwidth(2 << 30, 3)
    l2_bytes = backend.l2_bytes or (126 << 20)
    bvh_bytes = backend.bvh_bytes
    hbm_bound = bvh_bytes > l2_bytes
    compulsory = n_rays * (24 + out_bytes(mask)) + bvh_bytes
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(args.config)
        except Exception:
            traffic = None
    dram = traffic.get("dram_bytes") if isinstance(traffic, dict) else traffic
    peak = hbm_peak if hbm_bound else l2_gbs
    roofline = {
        "bound": "hbm" if hbm_bound else "l2", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": dram,
        "peak_source": peak_src if hbm_bound else
        "L2 read bandwidth measured in this run (prt_b200_read_bandwidth: 32 MiB re-read 50x; "
        "MEASURED_PEAKS.json has no L2 figure)",
        "kernel": "prt::k_trace<mask=%d,SoA,%s>" % (mask, "4-wide nodes" if hbm_bound else "binary nodes"),
        "kernel_ms": kern_ms, "step_ms_incl_ray_reordering": float(np.mean(step_ms)),
        "bytes_per_ray": bytes_per_ray, "nodes_per_ray": nodes_per_ray,
        "tris_per_ray": tris_per_ray, "l2_read_gbs_measured": l2_gbs,
        "hbm_read_gbs_measured": hbm_read_gbs, "frac_of_l2": achieved / l2_gbs,
        "frac_of_hbm": achieved / hbm_peak, "bvh_bytes": int(bvh_bytes), "l2_bytes": int(l2_bytes),
        "compulsory_bytes": compulsory,
        "dram_frac_of_hbm": (dram / (kern_ms * 1e-3) / 1e9 / hbm_peak) if dram else None,
        "ncu": traffic if isinstance(traffic, dict) else None,
        "note": ("BVH (%.1f MB) %s L2 (%.0f MB): %s" % (
            bvh_bytes / 1e6, "exceeds" if hbm_bound else "fits", l2_bytes / 1e6,
            "fetches that miss L1/L2 go to HBM; `traffic` is the ncu-measured DRAM bytes per launch, "
            "`dram_frac_of_hbm` the fraction of the HBM roof that real traffic amounts to" if hbm_bound else
            "fetched bytes are served by L1/L2, the L2 read bandwidth is the roof; DRAM traffic is "
            "compulsory only")),
    }
    # algorithmic HBM bytes per triangle: bounds 36 R; morton 36 R + 12 W; histograms
    # 8 R; P sort passes x (12 R + 12 W); hierarchy kernel: 4 (index) + 36 (triangle) + 16 (keys) R,
    # 64 W (record), 64 W (node halves), 24 R (sibling box), 8 (bound exchange)
    passes = 4 if n_tris <= (1 << 16) else (5 if n_tris <= (1 << 24) else 6)  # build.cu: morton_bits_for
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
             "frac_of_hbm": n_tris * build_bytes_per_tri / (build_ms_mean * 1e-3) / 1e9 / hbm_peak}

    per_mask = None
    if args.per_mask and world == 1:
        per_mask = {}
        uv_, t_, pid_ = torch.empty(n_rays, 2, device=dev), torch.empty(n_rays, device=dev), \
            torch.empty(n_rays, dtype=torch.int32, device=dev)
        p_, v_ = torch.empty(n_rays, 3, device=dev), torch.empty(n_rays, dtype=torch.uint8, device=dev)
        for combo in hitreg.TAG_COMBOS:
            m = hitreg.mask_of(combo)
            o = dict(uv=uv_.data_ptr() if m & 1 else 0, t=t_.data_ptr() if m & 2 else 0,
                     pid=pid_.data_ptr() if m & 4 else 0, p=p_.data_ptr() if m & 8 else 0,
                     valid=v_.data_ptr() if m & 16 else 0)
            ts = []
            for k in range(6):
                l2_flush()
                ms = backend.trace_dev(d_rays.data_ptr(), n_rays, m, **o)
                if k >= 2:
                    ts.append(ms)
            per_mask["_".join(combo)] = n_rays / (np.mean(ts) * 1e-3) / 1e6
        del uv_, t_, pid_, p_, v_

    # ---- parity against the unmodified reference on the fixed sample, and its throughput
    parity_rep, cpu_line = None, None
    if world == 1 and cpu is not None:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import parity as P
            from oracle import Oracle
            idx = sample_indices(n_rays)
            if dynamic:
                backend.set_tris_dev(d_tris.data_ptr(), n_tris)
            backend.trace_dev(d_rays.data_ptr(), n_rays, mask, **outs)
            ii = torch.from_numpy(idx).to(dev)
            got = {}
            if mask & 2:
                got["t"] = t[ii].cpu().numpy()
                got["valid"] = np.isfinite(got["t"])
            if mask & 4:
                got["pid"] = pid[ii].cpu().numpy().view(np.uint32)
            if mask & 1:
                got["u"], got["v"] = (x.cpu().numpy() for x in uv[ii].unbind(1))
            if mask & 8:
                got["px"], got["py"], got["pz"] = (x.cpu().numpy() for x in p[ii].unbind(1))
            if mask & 16:
                got["valid"] = valid[ii].cpu().numpy().astype(bool)
            cpu.wait()
            sample = np.ascontiguousarray(rays[idx])
            ref_hits, _ = cpu.trace(sample, 31)
            ref = P.from_structured(ref_hits)
            rep = P.compare(ref, got, tris, sample, Oracle())
            parity_rep = {"checked_rays": int(len(idx)),
                          "sample": "first 10^6 rays + every %d-th ray" % max(1, n_rays // 1_000_000)
                          if len(idx) < n_rays else "all rays",
                          "against": "unmodified reference CPU backend (oracle/_ref), full-size scene",
                          "valid_mismatch": rep["valid_mismatch"], "pid_mismatch": rep.get("pid_mismatch"),
                          "pid_ties": rep.get("pid_ties"), "t_maxrel": rep.get("t_maxrel"),
                          "t_bitexact": rep.get("t_bitexact"), "miss_t_not_inf": rep.get("miss_t_not_inf"),
                          "u_maxabs": rep.get("u_maxabs"), "v_maxabs": rep.get("v_maxabs"),
                          "p_maxrel": rep.get("p_maxrel"), "n_valid": rep["n_valid"]}
            cpu_line = cpu.timed(rays, mask, budget_s=20.0)
        except Exception as e:  # noqa: BLE001
            cpu_line = {"unavailable": repr(e)}
    elif world == 1 and not args.no_cpu_baseline:
        cpu_line = {"unavailable": cpu_err}

    # ---- opt-in watertight mode beside the default (reference arithmetic): its throughput and how
    # many rays of this batch it answers differently
    watertight = None
    if world == 1 and mask & 2 and not args.no_extras:
        if dynamic:  # compare on one and the same frame
            backend.set_tris_dev(d_tris.data_ptr(), n_tris)
        backend.trace_dev(d_rays.data_ptr(), n_rays, mask, **outs)
        t_def = t.clone()
        backend.set_triangle_test(1)
        # same tree state as the timed steps: optimised for static scenes, plain for the dynamic one
        backend.set_tree_optimisation(0 if tree_depth == 0 else 1, 2)
        backend.set_tris_dev(d_tris.data_ptr(), n_tris)
        ts = []
        for k in range(5):
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
        del t_def, rel, both, hit_def, hit_wt
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

    line = {
        "metric": "nearest_hits throughput", "value": value, "unit": "Mrays/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": config_dict(w, args),
        "rays_per_gpu": n_rays, "input_generation_s": gen_s,
        "build_mtris_s": build["mtris_s"], "build": build,
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "roofline": roofline, "parity": parity_rep, "cpu_baseline": cpu_line,
        "cpu_baseline_embree": embree_status() if world == 1 else None,
        "replicas": replicas, "watertight": watertight, "dynamic_without_reuse": dynamic_plain,
        "exotic_rays": int(backend.exotic_rays), "sorted_batches": int(backend.sorted_batches),
        "wall_s_timed_region": wall, "device": backend.device_name(),
    }
    if per_mask:
        line["per_mask_mrays_s"] = per_mask
    print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
