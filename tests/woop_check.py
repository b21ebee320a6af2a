"""TEST INFRASTRUCTURE ONLY: an independent numpy restatement of the opt-in watertight mode
(portablert_b200/csrc/prt_math.cuh: woop_watertight + slab_cons), brute force over rays x triangles.

numpy's float32 operators round to nearest and never fuse, exactly like the *_rn intrinsics of the
kernels, so the comparison is bit-for-bit.  Per ray the answer is the minimum t over every triangle
whose own AABB passes slab_cons and which woop_watertight accepts (lowest primitive id among equal
t) -- the same "per-triangle rule" shape the reference has (bvh.hpp:237-247), with the two
predicates swapped for their watertight counterparts."""
import numpy as np

F = np.float32


def brute(tris, rays, block=256):
    tris = np.ascontiguousarray(tris, F).reshape(-1, 3, 3)
    rays = np.ascontiguousarray(rays, F).reshape(-1, 6)
    R, N = len(rays), len(tris)
    out = {"t": np.full(R, np.inf, F), "u": np.zeros(R, F), "v": np.zeros(R, F),
           "pid": np.full(R, 0xFFFFFFFF, np.uint32)}
    lo, hi = tris.min(1), tris.max(1)
    with np.errstate(all="ignore"):
        for s in range(0, R, block):
            o, d = rays[s:s + block, None, :3], rays[s:s + block, None, 3:]
            n = o.shape[0]
            ad = np.abs(d[:, 0])
            kz = np.where((ad[:, 0] >= ad[:, 1]) & (ad[:, 0] >= ad[:, 2]), 0,
                          np.where(ad[:, 1] >= ad[:, 2], 1, 2))
            kx = np.where(kz == 2, 0, kz + 1)
            ky = np.where(kx == 2, 0, kx + 1)
            dz = d[np.arange(n), 0, kz]
            swap = dz < 0
            kx, ky = np.where(swap, ky, kx), np.where(swap, kx, ky)
            Sx = (d[np.arange(n), 0, kx] / dz)[:, None]
            Sy = (d[np.arange(n), 0, ky] / dz)[:, None]
            Sz = (F(1.0) / dz)[:, None]

            def comp(P, k):  # P: (n, N, 3)
                return np.take_along_axis(P, np.broadcast_to(k[:, None, None], (n, N, 1)), 2)[..., 0]
            A, B, C = tris[None, :, 0] - o, tris[None, :, 1] - o, tris[None, :, 2] - o
            Akz, Bkz, Ckz = comp(A, kz), comp(B, kz), comp(C, kz)
            Ax, Ay = comp(A, kx) - Sx * Akz, comp(A, ky) - Sy * Akz
            Bx, By = comp(B, kx) - Sx * Bkz, comp(B, ky) - Sy * Bkz
            Cx, Cy = comp(C, kx) - Sx * Ckz, comp(C, ky) - Sy * Ckz
            U, V, W = Cx * By - Cy * Bx, Ax * Cy - Ay * Cx, Bx * Ay - By * Ax
            z = (U == 0) | (V == 0) | (W == 0)
            D = np.float64
            U = np.where(z, (Cx.astype(D) * By - Cy.astype(D) * Bx).astype(F), U)
            V = np.where(z, (Ax.astype(D) * Cy - Ay.astype(D) * Cx).astype(F), V)
            W = np.where(z, (Bx.astype(D) * Ay - By.astype(D) * Ax).astype(F), W)
            rej = ((U < 0) | (V < 0) | (W < 0)) & ((U > 0) | (V > 0) | (W > 0))
            det = (U + V) + W
            rej |= det == 0
            T = (U * (Sz * Akz) + V * (Sz * Bkz)) + W * (Sz * Ckz)
            inv = F(1.0) / det
            t, u, v = T * inv, V * inv, W * inv
            # slab_cons on the triangle's own box
            idir = F(1.0) / d
            tl, th = (lo[None] - o) * idir, (hi[None] - o) * idir
            nan = np.isnan(tl) | np.isnan(th)  # 0 * inf: that axis does not bound the interval
            near = np.where(nan, -np.inf, np.fmin(tl, th)).astype(F)
            far = np.where(nan, np.inf, np.fmax(tl, th)).astype(F)
            tmin, tmax = near.max(-1), far.min(-1)
            eps = F(4.76837158e-7)
            tmin = tmin - np.fmin(np.abs(tmin), F(3.0e38)) * eps
            tmax = tmax + np.fmin(np.abs(tmax), F(3.0e38)) * eps
            box = ~(tmax < 0) & ~(tmin > tmax)
            ok = ~rej & box & (t < np.inf)  # NaN and +inf never update t_near (bvh.hpp:247)
            tt = np.where(ok, t, np.inf).astype(F)
            best = tt.min(1)
            # lowest primitive id among equal t: argmax of the first True
            first = (tt == best[:, None]).argmax(1)
            hit = best < np.inf
            idx = np.arange(n)
            out["t"][s:s + n] = best
            out["u"][s:s + n] = np.where(hit, u[idx, first], 0)
            out["v"][s:s + n] = np.where(hit, v[idx, first], 0)
            out["pid"][s:s + n] = np.where(hit, first, 0xFFFFFFFF)
    out["valid"] = out["t"] < np.inf
    return out
