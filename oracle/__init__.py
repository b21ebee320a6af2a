"""TEST INFRASTRUCTURE ONLY -- ctypes loaders for the CPU oracle.

Allowed importers: tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` /
``--impl reference`` legs.  The product package ``portablert_b200`` never imports this module.

* ``Oracle``      -- oracle/liboracle.so, the C restatement in oracle/oracle.c.
* ``Reference``   -- oracle/_ref/libprt_ref.so, the unmodified reference CPU backend
                     (include/portableRT/intersect_cpu.hpp + bvh.hpp + core.hpp) behind
                     oracle/ref_shim.cpp.  Built here from /root/reference; travels prebuilt to
                     the GPU box.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_ORACLE = os.path.join(HERE, "liboracle.so")
LIB_REF = os.path.join(HERE, "_ref", "libprt_ref.so")

FIELDS = ("t", "u", "v", "px", "py", "pz", "pid", "valid")


def build(quiet: bool = True) -> None:
    """(Re)build liboracle.so and, when /root/reference is present, oracle/_ref."""
    subprocess.run(["make", "-C", HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


class _Out(C.Structure):
    _fields_ = [("t", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p), ("px", C.c_void_p),
                ("py", C.c_void_p), ("pz", C.c_void_p), ("pid", C.c_void_p),
                ("valid", C.c_void_p)]


def _f32(a, cols):
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1, cols)
    return a


def _alloc(n):
    out = {k: np.empty(n, np.float32) for k in ("t", "u", "v", "px", "py", "pz")}
    out["pid"] = np.empty(n, np.uint32)
    out["valid"] = np.empty(n, np.uint8)
    st = _Out(*[out[k].ctypes.data for k in FIELDS])
    return out, st


class Oracle:
    """The C restatement.  ``build(tris)`` then ``trace(rays)``; or the topology-free ``brute``."""

    def __init__(self):
        if not os.path.exists(LIB_ORACLE):
            build()
        L = C.CDLL(LIB_ORACLE)
        L.oracle_build.restype = C.c_void_p
        L.oracle_build.argtypes = [C.c_void_p, C.c_uint64]
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_node_count.restype = C.c_uint64
        L.oracle_node_count.argtypes = [C.c_void_p]
        L.oracle_depth.restype = C.c_uint32
        L.oracle_depth.argtypes = [C.c_void_p]
        L.oracle_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(_Out),
                                   C.c_void_p, C.c_int]
        L.oracle_brute.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                   C.POINTER(_Out), C.c_int]
        L.oracle_candidate.argtypes = [C.c_void_p, C.c_void_p] + [C.POINTER(C.c_float)] * 3
        self.L = L
        self.h = None
        self.threads = os.cpu_count() or 1

    def build(self, tris):
        tris = _f32(tris, 9)
        self.free()
        self.h = self.L.oracle_build(tris.ctypes.data, len(tris))
        self.n_tris = len(tris)
        return self

    def free(self):
        if self.h:
            self.L.oracle_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    @property
    def node_count(self):
        return self.L.oracle_node_count(self.h)

    @property
    def depth(self):
        return self.L.oracle_depth(self.h)

    def trace(self, rays, visits=False, threads=None):
        rays = _f32(rays, 6)
        out, st = _alloc(len(rays))
        vis = np.zeros((len(rays), 2), np.uint32) if visits else None
        rc = self.L.oracle_trace(self.h, rays.ctypes.data, len(rays), C.byref(st),
                                 vis.ctypes.data if visits else None, threads or self.threads)
        if rc:
            raise RuntimeError("oracle traversal stack overflow (reference would overflow too)")
        out["valid"] = out["valid"].astype(bool)
        if visits:
            out["visits"] = vis
        return out

    def brute(self, tris, rays, threads=None):
        tris = _f32(tris, 9)
        rays = _f32(rays, 6)
        out, st = _alloc(len(rays))
        self.L.oracle_brute(tris.ctypes.data, len(tris), rays.ctypes.data, len(rays),
                            C.byref(st), threads or self.threads)
        out["valid"] = out["valid"].astype(bool)
        return out

    def candidate(self, tri9, ray6):
        """(accepted, t, u, v) of ONE triangle under the leaf rule (box test + Moeller-Trumbore)."""
        tri9 = np.ascontiguousarray(tri9, np.float32)
        ray6 = np.ascontiguousarray(ray6, np.float32)
        t, u, v = C.c_float(), C.c_float(), C.c_float()
        ok = self.L.oracle_candidate(tri9.ctypes.data, ray6.ctypes.data, C.byref(t), C.byref(u),
                                     C.byref(v))
        return bool(ok), np.float32(t.value), np.float32(u.value), np.float32(v.value)


class _Layout(C.Structure):
    _fields_ = [("stride", C.c_uint32)] + [(n, C.c_int32) for n in
                                           ("off_u", "off_v", "off_t", "off_pid", "off_valid",
                                            "off_px", "off_py", "off_pz")]


class Reference:
    """The unmodified reference CPU backend (all logical cores, intersect_cpu.hpp:26)."""

    def __init__(self):
        if not os.path.exists(LIB_REF):
            build()
        if not os.path.exists(LIB_REF):
            raise FileNotFoundError(
                f"{LIB_REF} missing: it is built from /root/reference by `make -C oracle` in the "
                "build container and travels prebuilt to the GPU box")
        L = C.CDLL(LIB_REF)
        L.ref_set_tris.restype = C.c_double
        L.ref_set_tris.argtypes = [C.c_void_p, C.c_uint64]
        L.ref_nearest_hits.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p,
                                       C.POINTER(C.c_double)]
        L.ref_nearest_hits_default.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.ref_layout.argtypes = [C.c_uint32, C.POINTER(_Layout)]
        L.ref_device_name.argtypes = [C.c_char_p, C.c_size_t]
        self.L = L
        self.threads = L.ref_hw_threads()
        self.last_trace_s = 0.0
        self.last_build_s = 0.0

    @staticmethod
    def available() -> bool:
        return os.path.exists(LIB_REF) or os.path.isdir("/root/reference/include")

    def device_name(self) -> str:
        buf = C.create_string_buffer(256)
        self.L.ref_device_name(buf, 256)
        return buf.value.decode()

    def layout(self, mask: int):
        l = _Layout()
        if self.L.ref_layout(mask, C.byref(l)):
            raise ValueError(f"bad mask {mask}")
        return (l.stride, l.off_u, l.off_v, l.off_t, l.off_pid, l.off_valid, l.off_px, l.off_py,
                l.off_pz)

    def set_tris(self, tris) -> float:
        tris = _f32(tris, 9)
        self.last_build_s = self.L.ref_set_tris(tris.ctypes.data, len(tris))
        return self.last_build_s

    def nearest_hits(self, rays, mask: int = 31, keep=True):
        """-> numpy structured array with the reference's own AoS layout for this tag mask."""
        from portablert_b200 import hitreg  # layout helper only (pure python)
        rays = _f32(rays, 6)
        dt = hitreg.dtype(mask)
        assert dt.itemsize == self.layout(mask)[0]
        hits = np.zeros(len(rays), dt) if keep else None
        s = C.c_double()
        rc = self.L.ref_nearest_hits(rays.ctypes.data, len(rays), mask,
                                     hits.ctypes.data if keep else None, C.byref(s))
        if rc:
            raise ValueError(f"bad mask {mask}")
        self.last_trace_s = s.value
        return hits
