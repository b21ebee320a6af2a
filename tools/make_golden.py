#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE
CPU backend (oracle/_ref/libprt_ref.so, built from /root/reference by oracle/Makefile).

Run in the build container (the only place /root/reference exists):
    python tools/make_golden.py
The fixtures travel to the GPU box; nothing there reads /root/reference.

Contents
  kat.npz          the reference's own known-answer rays (examples/validation/main.cpp:68-110,
                   README.md:144-165) plus the semantic probes of SURVEY.md section 9
  bunny.npz        examples/common/bunny.obj as (4968,9) float32 + the 128x128 validation-camera
                   rays (examples/validation/main.cpp:172-198) + the reference's FullHitReg output
  bunny_mask.npz   the 1024x1024 `valid` mask of examples/validation (the bunny.png the reference
                   checkout lacks), bit-packed
  c1.npz           config C1: every 16th of the 1 M rays, reference t
  c2.npz           config C2: every 53rd of the 1920x1080 rays on the 69 192-triangle blob, all tags
  soup.npz         random overlapping triangle soup with negative-t hits, all tags
  interior.npz     C3-style interior (8 000 triangles: large walls + small detail), 192x108 primary rays
  heightfield.npz  C5-style height field (frame 3, 4 800 triangles), 192x108 primary rays
(`python tools/make_golden.py interior heightfield` regenerates only the named ones)
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import Reference  # noqa: E402
from portablert_b200 import hitreg, scenes  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
REF_ROOT = "/root/reference"


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def full(ref, rays):
    return ref.nearest_hits(rays, hitreg.ALL)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = Reference()

    # ---- KATs and semantic probes
    cases = []

    def case(name, tris, ray):
        cases.append((name, np.asarray(tris, np.float32).reshape(-1, 9), np.asarray(ray, np.float32)))

    T = scenes.KAT_TRI
    case("kat_hit", T, [0.1, 0, -1, 0, 0, 1])            # validation/main.cpp:70-72 -> t=1,u=.3,v=.5
    case("kat_miss", T, [-2, 0, -1, 0, 0, 1])            # validation/main.cpp:74-76
    case("readme_hit", T, [0, 0, -1, 0, 0, 1])           # README.md:150-152
    case("neg_t", [[-1, -1, -1, 1, -1, -1, 0, 1, 1]], [0, -0.5, 0.5, 0, 0, 1])  # t=-1
    case("behind_flat", T, [0.1, 0, 1, 0, 0, 1])         # flat triangle behind: box tmax<0 -> miss
    case("edge", T, [0, -1, -1, 0, 0, 1])                # hit, u=.5 v=-0
    case("corner", T, [-1, -1, -1, 0, 0, 1])             # hit u=-0 v=-0
    case("apex", T, [0, 1, -1, 0, 0, 1])                 # miss
    case("parallel", T, [0, 0, 0, 1, 0, 0])              # det == 0 -> miss
    case("zero_dir", T, [0, 0, -1, 0, 0, 0])             # miss
    case("nonunit", T, [0, 0, -2.5, 0, 0, 2])            # t = 1.25
    case("far", T * np.float32(1e4), [100, 0, -1e6, 0, 0, 1])  # no tmax
    two = np.concatenate([T, T])
    case("tie_two_identical", two, [0.1, 0, -1, 0, 0, 1])  # reference picks pid 1
    fnn = np.concatenate([T + np.float32([0, 0, 1] * 3), T, T])
    case("tie_far_near_near", fnn, [0.1, 0, -1, 0, 0, 1])  # reference picks pid 2
    names, tri_list, ray_list, hit_list = [], [], [], []
    for name, tris, ray in cases:
        ref.set_tris(tris)
        h = full(ref, ray[None])
        names.append(name)
        tri_list.append(tris)
        ray_list.append(ray)
        hit_list.append(h[0])
        print(f"{name:20s} {h[0]}")
    np.savez_compressed(os.path.join(OUT, "kat.npz"), names=np.array(names),
                        tri_offsets=np.cumsum([0] + [len(t) for t in tri_list]),
                        tris=np.concatenate(tri_list), rays=np.stack(ray_list),
                        hits=np.array(hit_list, dtype=hitreg.dtype(hitreg.ALL)))

    # ---- bunny (real mesh)
    bunny = scenes.load_obj(os.path.join(REF_ROOT, "examples", "common", "bunny.obj"))
    assert bunny.shape == (4968, 9)
    ref.set_tris(bunny)
    rays = scenes.pinhole_rays(128, 128, cam=(0, 0, -0.5), sensor=0.05, dist=0.05, normalise=False)
    np.savez_compressed(os.path.join(OUT, "bunny.npz"), tris=bunny, rays=rays, hits=full(ref, rays))
    rays = scenes.pinhole_rays(1024, 1024, cam=(0, 0, -0.5), sensor=0.05, dist=0.05, normalise=False)
    mask = ref.nearest_hits(rays, hitreg.VALID)["valid"]
    print("bunny 1024^2 coverage", mask.mean())
    np.savez_compressed(os.path.join(OUT, "bunny_mask.npz"), mask=np.packbits(mask),
                        rays_sha256=digest(rays))

    # ---- C1
    rays = scenes.c1_rays()
    ref.set_tris(scenes.KAT_TRI)
    t = ref.nearest_hits(rays, hitreg.T)["t"]
    np.savez_compressed(os.path.join(OUT, "c1.npz"), stride=16, t=t[::16], rays_sha256=digest(rays),
                        hit_fraction=np.isfinite(t).mean())

    # ---- C2
    tris = scenes.blob()
    rays = scenes.pinhole_rays(1920, 1080)
    ref.set_tris(tris)
    sel = np.arange(0, len(rays), 53)
    np.savez_compressed(os.path.join(OUT, "c2.npz"), stride=53, hits=full(ref, rays[sel]),
                        tris_sha256=digest(tris), rays_sha256=digest(rays))

    # ---- soup with negative-t hits
    g = np.random.default_rng(11)
    soup = (g.random((4000, 3, 3), dtype=np.float32) * 2 - 1)
    soup = (soup[:, :1] + (soup - soup[:, :1]) * np.float32(0.25)).reshape(-1, 9).astype(np.float32)
    rays = scenes.incoherent_rays(30000, [-1.0] * 3, [1.0] * 3, seed=12)
    ref.set_tris(soup)
    h = full(ref, rays)
    print("soup: valid", h["valid"].mean(), "negative t", (h["t"] < 0).sum())
    np.savez_compressed(os.path.join(OUT, "soup.npz"), tris=soup, rays=rays, hits=h)

    extra_scenes(ref)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


def extra_scenes(ref, only=None):
    """C3- and C5-style scenes at fixture size: triangles, rays and the reference's full records"""
    todo = {
        "interior": (scenes.interior(8000), scenes.camera_rays(192, 108, (2, 6, 3), (28, 4, 15))),
        "heightfield": (scenes.heightfield(frame=3, nx=60, nz=40),
                        scenes.camera_rays(192, 108, (10, 6, -4), (10, 0, 5))),
    }
    for name, (tris, rays) in todo.items():
        if only and name not in only:
            continue
        ref.set_tris(tris)
        h = full(ref, rays)
        print(name, len(tris), "tris, valid", h["valid"].mean())
        np.savez_compressed(os.path.join(OUT, name + ".npz"), tris=tris, rays=rays, hits=h)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        extra_scenes(Reference(), only=set(sys.argv[1:]))
    else:
        main()
