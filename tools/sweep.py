#!/usr/bin/env python
"""Kernel-tuning sweep on one GPU: device-timed traversal of a workload under several settings of
the PRT_B200_* knobs (read at context creation), one scene upload, one process.

    python tools/sweep.py --config c2 --set LEAF_VOTES=1,4,8,16 --set REFILL=16,24 [--steps 6]

Prints one JSON line per combination: ms per step (reordering + traversal), the traversal kernel
alone, Mrays/s.  Inputs are cached under /tmp so that several invocations (e.g. one per library
variant, PRT_B200_LIB=...) do not regenerate them.
"""
import argparse
import itertools
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--set", action="append", default=[])
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--tag", default="")
    ap.add_argument("--repeat", type=int, default=1, help="trace the batch tiled this many times")
    ap.add_argument("--passes", default="2", help="treelet passes, comma-separated list to sweep")
    args = ap.parse_args()
    import torch

    import bench
    import portablert_b200 as prt
    dev = torch.device("cuda", 0)
    cache = f"/tmp/prt_sweep_{args.config}_{os.environ.get('PRT_BENCH_C4_RAYS', 'full')}"
    if os.path.exists(cache + "_rays.npy"):
        tris, rays = np.load(cache + "_tris.npy"), np.load(cache + "_rays.npy", mmap_mode="r")
        mask = int(open(cache + "_mask").read())
    else:
        def tracer(t, r):
            b = prt.CUDABackend(device=0)
            b.init()
            b.set_tris(t)
            h = b.nearest_hits(r)
            b.shutdown()
            return h
        w = bench.workload(args.config, tracer=tracer)
        tris, rays, mask = w["tris"], w["rays"], w["mask"]
        np.save(cache + "_tris.npy", tris)
        np.save(cache + "_rays.npy", rays)
        open(cache + "_mask", "w").write(str(mask))
    if args.repeat > 1:
        rays = np.tile(np.asarray(rays), (args.repeat, 1))
    d_tris = torch.from_numpy(np.ascontiguousarray(tris)).to(dev)
    d_rays = torch.from_numpy(np.ascontiguousarray(rays)).to(dev)
    n = len(rays)
    t = torch.empty(n, device=dev)
    pid = torch.empty(n, dtype=torch.int32, device=dev)
    uv = torch.empty(n, 2, device=dev)
    p = torch.empty(n, 3, device=dev)
    valid = torch.empty(n, dtype=torch.uint8, device=dev)
    outs = dict(uv=uv.data_ptr() if mask & 1 else 0, t=t.data_ptr() if mask & 2 else 0,
                pid=pid.data_ptr() if mask & 4 else 0, p=p.data_ptr() if mask & 8 else 0,
                valid=valid.data_ptr() if mask & 16 else 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    names, values = ["passes"], [args.passes.split(",")]
    for s in args.set:
        k, v = s.split("=")
        names.append(k)
        values.append(v.split(","))
    ref = None
    for combo in itertools.product(*values):
        for k, v in zip(names[1:], combo[1:]):
            os.environ["PRT_B200_" + k] = v
        b = prt.CUDABackend(device=0)
        b.init()
        b.set_tree_optimisation(1, int(combo[0]))  # optimised inside set_tris: steady state of a static scene
        build = b.set_tris_dev(d_tris.data_ptr(), len(tris))
        ms, km = [], []
        for i in range(2 + args.steps):
            flush.zero_()
            torch.cuda.synchronize()
            x = b.trace_dev(d_rays.data_ptr(), n, mask, **outs)
            if i >= 2:
                ms.append(x)
                km.append(b.last_kernel_ms if not os.environ.get('PRT_B200_LIB_OLD_ABI') else x)
        digest = bench.checksum(t if mask & 2 else None, pid if mask & 4 else None,
                                valid if (mask & 16) and not (mask & 2) else None)
        if ref is None:
            ref = digest
        line = {"tag": args.tag, "config": args.config, "repeat": args.repeat, **dict(zip(names, combo)),
                "ms": round(float(np.mean(ms)), 4), "kernel_ms": round(float(np.mean(km)), 4),
                "mrays_s": round(n / float(np.mean(ms)) / 1e3, 1), "build_opt_ms": round(build, 3),
                "same_result": digest == ref}
        print(json.dumps(line), flush=True)
        b.shutdown()


if __name__ == "__main__":
    main()
