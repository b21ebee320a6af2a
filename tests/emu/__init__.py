"""TEST INFRASTRUCTURE ONLY: ctypes loader for tests/emu/libprt_emu.so (see emu.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libprt_emu.so")
SRC = os.path.join(HERE, "emu.cpp")
CSRC = os.path.join(HERE, "..", "..", "portablert_b200", "csrc")


def build():
    deps = [SRC] + [os.path.join(CSRC, h) for h in ("prt_math.cuh", "prt_traverse.cuh",
                                                      "prt_treelet.cuh")]
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", LIB,
                    SRC], check=True)


class Emu:
    def __init__(self):
        build()
        L = C.CDLL(LIB)
        L.emu_build.restype = C.c_void_p
        L.emu_build.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int]
        L.emu_load.restype = C.c_void_p
        L.emu_load.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int32]
        L.emu_free.argtypes = [C.c_void_p]
        L.emu_refit.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.emu_check_split_search.restype = C.c_int
        L.emu_check_split_search.argtypes = [C.c_void_p]
        L.emu_treelet.restype = C.c_int32
        L.emu_treelet.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
        L.emu_num_nodes.restype = C.c_uint64
        L.emu_num_nodes.argtypes = [C.c_void_p]
        L.emu_download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.emu_download_wide.argtypes = [C.c_void_p, C.c_void_p]
        L.emu_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_float, C.c_float,
                                C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 8
        self.L = L
        self.h = None

    def build(self, tris, bits=16, watertight=False):
        tris = np.ascontiguousarray(tris, np.float32).reshape(-1, 9)
        self.free()
        self.n = len(tris)
        self.h = self.L.emu_build(tris.ctypes.data, len(tris), bits, int(watertight))
        return self

    def load(self, nodes, tris, root=0):
        self.free()
        nodes = np.ascontiguousarray(nodes)
        tris = np.ascontiguousarray(tris)
        self.n = tris.nbytes // 64
        self.h = self.L.emu_load(nodes.ctypes.data, nodes.nbytes // 64, tris.ctypes.data, self.n, root)
        return self

    def treelet(self, passes=1, strict=False):
        """opt-in SAH optimisation (prt_treelet.cuh) -> (tree height, treelets changed in last pass)"""
        ch = C.c_uint64(0)
        d = self.L.emu_treelet(self.h, int(passes), int(strict), C.byref(ch))
        return int(d), int(ch.value)

    def check_split_search(self, copt):
        copt = np.ascontiguousarray(copt, np.float32)
        assert copt.size == 128
        return int(self.L.emu_check_split_search(copt.ctypes.data))

    def refit(self, tris, watertight=False):
        """same topology, new vertices (stand-in for k_refit)"""
        tris = np.ascontiguousarray(tris, np.float32).reshape(-1, 9)
        assert len(tris) == self.n
        self.L.emu_refit(self.h, tris.ctypes.data, int(watertight))
        return self

    def free(self):
        if self.h:
            self.L.emu_free(self.h)
            self.h = None

    def download(self):
        nn = self.L.emu_num_nodes(self.h)
        nodes = np.zeros((nn, 16), np.float32)
        tris = np.zeros((self.n, 16), np.float32)
        self.L.emu_download(self.h, nodes.ctypes.data, tris.ctypes.data)
        return nodes, tris

    def download_wide(self):
        nn = self.L.emu_num_nodes(self.h)
        out = np.zeros((nn, 64), np.uint8)
        self.L.emu_download_wide(self.h, out.ctypes.data)
        return out

    def trace(self, rays, prune=1, slack_rel=1e-4, slack_ulps=64.0, anyhit=False, fast=True, wide=False,
              watertight=False):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = len(rays)
        o = {k: np.empty(n, np.float32) for k in ("t", "u", "v")}
        o["pid"] = np.empty(n, np.uint32)
        o["valid"] = np.empty(n, np.uint8)
        p = np.empty((n, 3), np.float32)
        cnt = np.zeros((n, 2), np.uint32)
        ff = np.zeros(n, np.uint8)
        self.L.emu_trace(self.h, rays.ctypes.data, n, int(prune), slack_rel, slack_ulps,
                         int(anyhit), int(fast), int(wide), int(watertight), o["t"].ctypes.data, o["u"].ctypes.data, o["v"].ctypes.data,
                         o["pid"].ctypes.data, o["valid"].ctypes.data, p.ctypes.data,
                         cnt.ctypes.data, ff.ctypes.data)
        o["fast"] = ff.astype(bool)
        o["valid"] = o["valid"].astype(bool)
        o["px"], o["py"], o["pz"] = (np.ascontiguousarray(p[:, k]) for k in range(3))
        o["counts"] = cnt
        return o
