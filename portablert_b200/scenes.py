"""Seeded synthetic scenes and ray batches for the five BASELINE.json configs.

Everything is float32 numpy: triangles are ``(N, 9)`` (v0 xyz, v1 xyz, v2 xyz -- the reference's
``Tri = std::array<float,9>``, core.hpp:24) and rays are ``(R, 6)`` (origin xyz, direction xyz --
``Ray``, core.hpp:19-22).  The same arrays feed the CPU oracle and the GPU backend.
"""
from __future__ import annotations

import numpy as np

F = np.float32


def _rng(seed):
    return np.random.Generator(np.random.MT19937(seed))


# --------------------------------------------------------------------------- C1
KAT_TRI = np.array([[-1, -1, 0, 1, -1, 0, 0, 1, 0]], F)  # examples/triangle/main.cpp:7


def c1_rays(n=1_000_000, seed=1):
    """Origins (x, y, -1), x,y ~ U[-2,2); first half d=(0,0,1), second half jittered+normalised."""
    g = _rng(seed)
    xy = g.uniform(-2.0, 2.0, size=(n, 2)).astype(F)
    rays = np.zeros((n, 6), F)
    rays[:, 0:2] = xy
    rays[:, 2] = -1.0
    rays[:, 5] = 1.0
    h = n // 2
    j = g.uniform(-0.25, 0.25, size=(n - h, 2)).astype(F)
    d = np.concatenate([j, np.ones((n - h, 1), F)], axis=1)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(F)
    rays[h:, 3:6] = d.astype(F)
    return rays


# --------------------------------------------------------------------------- meshes
def _grid_tris(P):
    """P: (a+1, b+1, 3) vertex grid -> (2ab, 9) triangles, two per quad."""
    p00 = P[:-1, :-1]
    p10 = P[1:, :-1]
    p01 = P[:-1, 1:]
    p11 = P[1:, 1:]
    t0 = np.concatenate([p00, p10, p11], axis=-1).reshape(-1, 9)
    t1 = np.concatenate([p00, p11, p01], axis=-1).reshape(-1, 9)
    out = np.empty((t0.shape[0] * 2, 9), F)
    out[0::2] = t0
    out[1::2] = t1
    return out


def blob(nu=186, nv=186, radius=0.1, bump=0.15, center=(0.0, 0.0, 0.0)):
    """Closed displaced UV sphere, 2*nu*nv triangles (69 192 at the default = config C2).
    The pole bands contain zero-area triangles on purpose (det == 0 path, core.hpp:42)."""
    th = np.linspace(0.0, np.pi, nv + 1)[:, None]
    ph = np.linspace(0.0, 2.0 * np.pi, nu + 1)[None, :]
    r = radius * (1.0 + bump * np.sin(3.0 * th) * np.cos(4.0 * ph)
                  + 0.5 * bump * np.sin(7.0 * th) * np.sin(5.0 * ph))
    x = r * np.sin(th) * np.cos(ph) + center[0]
    y = r * np.cos(th) * np.ones_like(ph) + center[1]
    z = r * np.sin(th) * np.sin(ph) + center[2]
    P = np.stack([x, y, z], axis=-1).astype(F)
    P[:, -1] = P[:, 0]  # close the seam exactly
    return _grid_tris(P)


def pinhole_rays(width, height, cam=(0.0, 0.0, -0.3), sensor=0.05, dist=0.05, normalise=True):
    """Camera of examples/bunny/main.cpp:101-124: rows top->bottom, sensor plane at cam.z+dist,
    unit directions (the validation example's non-unit variant: ``normalise=False`` divides by
    |sensor_pos| like examples/validation/main.cpp:189-194)."""
    ys = np.arange(height - 1, -1, -1, dtype=F)
    xs = np.arange(width, dtype=F)
    sx = (F(sensor) * (xs / F(width) - F(0.5))).astype(F)
    sy = (F(sensor) * (ys / F(height) - F(0.5))).astype(F)
    SX, SY = np.meshgrid(sx, sy)
    cam = np.asarray(cam, F)
    sp = np.stack([SX, SY, np.full_like(SX, cam[2] + F(dist))], axis=-1).astype(F)
    d = (sp - cam).astype(F)
    if normalise:
        ln = np.sqrt((d * d).sum(-1, keepdims=True)).astype(F)
    else:
        ln = np.sqrt((sp * sp).sum(-1, keepdims=True)).astype(F)
    d = (d / ln).astype(F)
    rays = np.empty((height * width, 6), F)
    rays[:, 0:3] = cam
    rays[:, 3:6] = d.reshape(-1, 3)
    return rays


def camera_rays(width, height, eye, target, up=(0, 1, 0), fov_deg=60.0):
    """Generic look-at pinhole camera, unit directions, rows top->bottom."""
    eye = np.asarray(eye, np.float64)
    f = np.asarray(target, np.float64) - eye
    f /= np.linalg.norm(f)
    r = np.cross(f, np.asarray(up, np.float64))
    r /= np.linalg.norm(r)
    u = np.cross(r, f)
    th = np.tan(np.radians(fov_deg) / 2)
    xs = ((np.arange(width) + 0.5) / width * 2 - 1) * th * width / height
    ys = (1 - (np.arange(height) + 0.5) / height * 2) * th
    X, Y = np.meshgrid(xs, ys)
    d = f[None, None] + X[..., None] * r[None, None] + Y[..., None] * u[None, None]
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    rays = np.empty((height * width, 6), F)
    rays[:, 0:3] = eye.astype(F)
    rays[:, 3:6] = d.reshape(-1, 3).astype(F)
    return rays


def _quad(p0, p1, p2, p3, nu=1, nv=1):
    """Tessellated parallelogram p0->p1 (u) / p0->p3 (v)."""
    p0, p1, p3 = (np.asarray(p, np.float64) for p in (p0, p1, p3))
    u = np.linspace(0, 1, nu + 1)[:, None, None]
    v = np.linspace(0, 1, nv + 1)[None, :, None]
    P = p0 + u * (p1 - p0) + v * (p3 - p0)
    return _grid_tris(P.astype(F))


def _box(lo, hi, n=1):
    lo = np.asarray(lo, np.float64)
    hi = np.asarray(hi, np.float64)
    x0, y0, z0 = lo
    x1, y1, z1 = hi
    q = [
        ((x0, y0, z0), (x1, y0, z0), None, (x0, y1, z0)),
        ((x0, y0, z1), (x1, y0, z1), None, (x0, y1, z1)),
        ((x0, y0, z0), (x0, y0, z1), None, (x0, y1, z0)),
        ((x1, y0, z0), (x1, y0, z1), None, (x1, y1, z0)),
        ((x0, y0, z0), (x1, y0, z0), None, (x0, y0, z1)),
        ((x0, y1, z0), (x1, y1, z0), None, (x0, y1, z1)),
    ]
    return np.concatenate([_quad(a, b, c, d, n, n) for a, b, c, d in q])


def interior(target_tris=262_144, seed=3):
    """Sponza-scale interior (config C3): 30 x 12 x 18 room made of a few huge wall quads, a grid
    of square columns, tessellated half-cylinder arches and finely tessellated draped curtains.
    Triangle edge lengths span > 1e3 (SAH stress like Sponza)."""
    g = _rng(seed)
    parts = [_box((0, 0, 0), (30, 12, 18), 1)]  # 12 huge triangles
    cols_x = np.linspace(4, 26, 8)
    cols_z = (5.0, 13.0)
    for cx in cols_x:
        for cz in cols_z:
            parts.append(_box((cx - 0.4, 0, cz - 0.4), (cx + 0.4, 8, cz + 0.4), 6))
    # arches between neighbouring columns: half cylinders along x
    for cz in cols_z:
        for a, b in zip(cols_x[:-1], cols_x[1:]):
            r = (b - a) / 2 - 0.4
            xc = (a + b) / 2
            ang = np.linspace(0, np.pi, 25)[:, None]
            zz = np.linspace(cz - 0.4, cz + 0.4, 5)[None, :]
            P = np.stack([xc + r * np.cos(ang) * np.ones_like(zz),
                          8 + r * 0.6 * np.sin(ang) * np.ones_like(zz),
                          np.ones_like(ang) * zz], axis=-1)
            parts.append(_grid_tris(P.astype(F)))
    fixed = sum(len(p) for p in parts)
    # curtains: height fields hanging between the column rows; take the rest of the budget
    n_curt = 6
    per = max(2, (target_tris - fixed) // n_curt)
    m = max(2, int(np.sqrt(per / 2)))
    for k in range(n_curt):
        x0 = 3.0 + 4.2 * k
        u = np.linspace(0, 1, m + 1)[:, None]
        v = np.linspace(0, 1, m + 1)[None, :]
        phase = g.uniform(0, 2 * np.pi)
        xx = x0 + 3.0 * u + 0 * v
        yy = 2.0 + 7.0 * v + 0 * u
        zz = 9.0 + 0.35 * np.sin(14 * np.pi * u + phase) * (1 - 0.6 * v) + 0.05 * np.cos(9 * v)
        parts.append(_grid_tris(np.stack([xx, yy, zz], -1).astype(F)))
    return np.concatenate(parts).astype(F)


def heightfield(frame=0, nx=1000, nz=500, amp=0.5):
    """Config C5: nx*nz*2 = 1 000 000 triangles, y = A sin(kx + 0.1 f) cos(kz), re-evaluated per frame."""
    x = np.linspace(0, 20, nx + 1)[:, None]
    z = np.linspace(0, 10, nz + 1)[None, :]
    y = amp * np.sin(2.0 * x + 0.1 * frame) * np.cos(2.0 * z)
    P = np.stack([x + 0 * z, y, z + 0 * x], -1).astype(F)
    return _grid_tris(P)


def sphere_field(n_spheres=10_000, nu=25, nv=20, extent=1000.0, seed=4):
    """Config C4: n_spheres copies of a 2*nu*nv-triangle noisy UV sphere (1 000 tris at the
    default), radius U[1,5), centres U[0,extent)^3, baked into one flat triangle list."""
    g = np.random.Generator(np.random.MT19937(seed))
    th = np.linspace(0.0, np.pi, nv + 1)[:, None]
    ph = np.linspace(0.0, 2.0 * np.pi, nu + 1)[None, :]
    unit = np.stack([np.sin(th) * np.cos(ph), np.cos(th) * np.ones_like(ph),
                     np.sin(th) * np.sin(ph)], -1)
    unit[:, -1] = unit[:, 0]
    base = _grid_tris(unit.astype(F)).reshape(-1, 3, 3)  # (T,3,3)
    T = base.shape[0]
    out = np.empty((n_spheres, T, 3, 3), F)
    rad = g.uniform(1.0, 5.0, n_spheres).astype(F)
    cen = g.uniform(0.0, extent, (n_spheres, 3)).astype(F)
    chunk = 500
    for s in range(0, n_spheres, chunk):
        e = min(n_spheres, s + chunk)
        noise = (1.0 + 0.05 * g.standard_normal((e - s, T, 3, 1))).astype(F)
        out[s:e] = base[None] * noise * rad[s:e, None, None, None] + cen[s:e, None, None, :]
    return out.reshape(-1, 9)


def incoherent_rays(n, lo, hi, seed=4):
    """Origins uniform in the box [lo,hi], directions uniform on the sphere (config C4)."""
    g = _rng(seed)
    lo = np.asarray(lo, F)
    hi = np.asarray(hi, F)
    rays = np.empty((n, 6), F)
    rays[:, 0:3] = (lo + (hi - lo) * g.random((n, 3), dtype=F)).astype(F)
    d = g.standard_normal((n, 3), dtype=F)
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), F(1e-20))
    rays[:, 3:6] = d
    return rays


RAY_BLOCK = 1 << 20


def incoherent_rays_range(first, count, lo, hi, seed=4):
    """Rays [first, first+count) of an unbounded incoherent batch (config C4: 100 M rays): every
    block of 2^20 rays has its own generator MT19937(seed * 1000003 + block), so any contiguous
    slice -- one rank's share of a strong-scaled batch -- is produced without generating the
    rays in front of it, and is identical to the same rows of the whole batch."""
    out = np.empty((count, 6), F)
    at = first
    end = first + count
    while at < end:
        b = at // RAY_BLOCK
        blk = incoherent_rays(RAY_BLOCK, lo, hi, seed=seed * 1000003 + b)
        a0 = at - b * RAY_BLOCK
        n = min(end - at, RAY_BLOCK - a0)
        out[at - first: at - first + n] = blk[a0: a0 + n]
        at += n
    return out


def _hash_u01(idx, salt):
    """Counter-based per-ray uniform in [0,1): a 32-bit integer hash of (ray index, salt)."""
    x = (idx.astype(np.uint64) * np.uint64(0x9E3779B1) + np.uint64(salt)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(13)
    x = (x * np.uint64(0xC2B2AE35)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    return (x >> np.uint64(8)).astype(np.float64) / float(1 << 24)


def bounce_rays(tris, rays, valid, pid, p, seed=7, offset=1e-3):
    """One-bounce diffuse rays for config C3: at each valid primary hit, origin = p + offset*n
    (n = geometric normal flipped toward the incoming ray; the offset keeps the oracle's
    negative-t self hits away), direction = cosine-weighted hemisphere sample
    from a counter-based hash of the ray index.  Returns (bounce_rays, index of parent ray)."""
    idx = np.nonzero(valid)[0]
    tv = tris[pid[idx]].astype(np.float64).reshape(-1, 3, 3)
    n = np.cross(tv[:, 1] - tv[:, 0], tv[:, 2] - tv[:, 0])
    ln = np.linalg.norm(n, axis=1, keepdims=True)
    ok = ln[:, 0] > 0
    idx, tv, n, ln = idx[ok], tv[ok], n[ok], ln[ok]
    n /= ln
    d_in = rays[idx, 3:6].astype(np.float64)
    flip = (n * d_in).sum(1) > 0
    n[flip] *= -1
    u1 = _hash_u01(idx, seed * 2 + 1)
    u2 = _hash_u01(idx, seed * 2 + 2)
    r = np.sqrt(u1)
    phi = 2 * np.pi * u2
    a = np.where(np.abs(n[:, :1]) > 0.9, np.array([[0.0, 1.0, 0.0]]), np.array([[1.0, 0.0, 0.0]]))
    t1 = np.cross(n, a)
    t1 /= np.linalg.norm(t1, axis=1, keepdims=True)
    t2 = np.cross(n, t1)
    d = (r * np.cos(phi))[:, None] * t1 + (r * np.sin(phi))[:, None] * t2 \
        + np.sqrt(np.maximum(0.0, 1 - u1))[:, None] * n
    out = np.empty((len(idx), 6), F)
    out[:, 0:3] = (p[idx].astype(np.float64) + offset * n).astype(F)
    out[:, 3:6] = d.astype(F)
    return out, idx


def load_obj(path):
    """Minimal OBJ reader ('v' and triangular 'f' lines only, like examples/common/bunny.obj)."""
    vs, fs = [], []
    with open(path) as fh:
        for line in fh:
            if line.startswith("v "):
                vs.append([float(x) for x in line.split()[1:4]])
            elif line.startswith("f "):
                ids = [int(tok.split("/")[0]) for tok in line.split()[1:]]
                if len(ids) == 3:  # the reference skips non-triangles (validation/main.cpp:148-149)
                    fs.append(ids)
    v = np.asarray(vs, F)
    f = np.asarray(fs, np.int64) - 1
    return v[f].reshape(-1, 9).astype(F)
