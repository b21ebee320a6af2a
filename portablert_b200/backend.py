"""Python host mirror of the reference's backend interface for the CUDA (B200) backend.

Same names, argument meaning and error behaviour as the reference's C++ API so that the parity
tests read like the reference's own examples:

    reference (C++)                                         here (Python)
    ------------------------------------------------------  ------------------------------------
    portableRT::all_backends() / available_backends()        all_backends() / available_backends()
      backend.hpp:42-46
    portableRT::select_backend(b)     src/backend.cpp:50-56  select_backend(b)
    b->set_tris(tris)                 backend.hpp:19         b.set_tris(tris)          (N,9) float32
    portableRT::nearest_hits<Tags...>(rays)                  nearest_hits(rays, "t", "valid")
      nearest_hits_impl.hpp:28-36
    portableRT::nearest_hits(rays) -> FullHitReg             nearest_hits(rays)
      backend.hpp:77-79
    b->nearest_hits<Tags...>(rays)    intersect_cpu.hpp:20   b.nearest_hits(rays, "t", "valid")

Hits come back as a numpy structured array whose dtype has exactly the byte layout of the
reference's ``HitReg<Tags...>`` (hitreg.py).  All computation happens in libprt_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import hitreg
from ._lib import HitLayout, SoaOut, TraceOpts, lib


class Backend:
    """backend.hpp:12-29."""

    def __init__(self, name: str):
        self._name = name

    def name(self) -> str:
        return self._name

    def set_tris(self, tris):
        raise NotImplementedError

    def is_available(self) -> bool:
        raise NotImplementedError

    def init(self):
        raise NotImplementedError

    def shutdown(self):
        raise NotImplementedError

    def device_name(self) -> str:
        raise NotImplementedError


def _layout_struct(mask: int) -> HitLayout:
    return HitLayout(*hitreg.layout_tuple(mask))


class CUDABackend(Backend):
    """The B200-native backend ("CUDA"), sibling of the reference's CPU/OptiX/HIP/SYCL/Embree ones."""

    def __init__(self, device: int | None = None, gpus: int | None = None):
        """device: CUDA device index (None: env PRT_B200_DEVICE / the first CC 10.x device).
        gpus: number of GPUs this one backend spreads over inside the library (None: env
        PRT_B200_GPUS, default 1) -- the scene is replicated over NVLink, host ray batches are cut
        into contiguous slices, hits come back in ray order (prt_b200_create_multi)."""
        super().__init__("CUDA")
        self._h = C.c_void_p()
        self._device = device
        self._gpus = gpus

    # --- Backend virtuals -----------------------------------------------------------------------
    def is_available(self) -> bool:
        return lib().prt_b200_device_count() > 0

    def init(self):
        """Re-entrant like select_backend() requires (it may shutdown()+init() the same object,
        examples/validation/main.cpp:242)."""
        if self._h:
            return
        if self._gpus is not None:
            rc = lib().prt_b200_create_multi(C.byref(self._h), int(self._gpus))
        else:
            dev = -1 if self._device is None else int(self._device)
            rc = lib().prt_b200_create(C.byref(self._h), dev)
        if rc:
            self._h = C.c_void_p()
            raise RuntimeError("CUDA backend init failed: " + lib().prt_b200_last_error(None).decode())
        prune = os.environ.get("PRT_B200_PRUNE")
        if prune is not None:
            self.set_trace_opts(prune=int(prune))

    def shutdown(self):
        if self._h:
            lib().prt_b200_destroy(self._h)
            self._h = C.c_void_p()

    def device_name(self) -> str:
        self._need()
        buf = C.create_string_buffer(256)
        lib().prt_b200_device_name(self._h, buf, 256)
        return buf.value.decode()

    def set_tris(self, tris):
        self._need()
        tris = np.ascontiguousarray(tris, np.float32).reshape(-1, 9)
        self._check(lib().prt_b200_set_tris(self._h, tris.ctypes.data, len(tris)))

    # --- nearest_hits<Tags...> ------------------------------------------------------------------
    def nearest_hits(self, rays, *tags, out=None):
        """Host rays (R,6) float32 -> structured array of HitReg<tags...> (all tags when none given).
        `out` (optional) is a preallocated result array, e.g. from pinned_empty(), to be reused."""
        self._need()
        mask = hitreg.mask_of(tags[0] if len(tags) == 1 and not isinstance(tags[0], str) else tags)
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        if out is None:
            hits = np.zeros(len(rays), hitreg.dtype(mask))
        else:
            hits = out
            if hits.dtype != hitreg.dtype(mask) or len(hits) != len(rays) or \
                    not hits.flags["C_CONTIGUOUS"]:
                raise ValueError("out must be a contiguous HitReg array of len(rays) for these tags")
        lay = _layout_struct(mask)
        self._check(lib().prt_b200_nearest_hits(self._h, rays.ctypes.data, len(rays), mask,
                                                C.byref(lay), hits.ctypes.data))
        return hits

    # --- device-resident entry points (device pointers as ints, e.g. torch.Tensor.data_ptr()) ---
    def set_tris_dev(self, d_tris_ptr: int, n_tris: int) -> float:
        self._need()
        ms = C.c_float()
        self._check(lib().prt_b200_set_tris_dev(self._h, d_tris_ptr, n_tris, C.byref(ms)))
        return ms.value

    def trace_dev(self, d_rays_ptr: int, n_rays: int, mask: int, uv=0, t=0, pid=0, p=0,
                  valid=0) -> float:
        self._need()
        out = SoaOut(uv or None, t or None, pid or None, p or None, valid or None)
        ms = C.c_float()
        self._check(lib().prt_b200_trace_dev(self._h, d_rays_ptr, n_rays, mask, C.byref(out),
                                             C.byref(ms)))
        return ms.value

    def trace_dev_aos(self, d_rays_ptr: int, n_rays: int, mask: int, d_hits_ptr: int) -> float:
        self._need()
        lay = _layout_struct(mask)
        ms = C.c_float()
        self._check(lib().prt_b200_trace_dev_aos(self._h, d_rays_ptr, n_rays, mask, C.byref(lay),
                                                 d_hits_ptr, C.byref(ms)))
        return ms.value

    def trace_count_dev(self, d_rays_ptr: int, n_rays: int, d_counts_ptr: int):
        self._need()
        self._check(lib().prt_b200_trace_count_dev(self._h, d_rays_ptr, n_rays, d_counts_ptr))

    def set_trace_opts(self, prune=1, slack_rel=1e-4, slack_ulps=64.0):
        self._need()
        o = TraceOpts(int(prune), float(slack_rel), float(slack_ulps))
        self._check(lib().prt_b200_set_trace_opts(self._h, C.byref(o)))

    def set_ray_sorting(self, mode: int):
        """0 never, 1 always, 2 automatic (default): reorder incoherent batches before traversal."""
        self._need()
        self._check(lib().prt_b200_set_ray_sorting(self._h, int(mode)))

    def set_tree_optimisation(self, mode: int = 3, passes: int = 2):
        """SAH optimisation of the LBVH by treelet restructuring (Karras & Aila 2013): mode 0 never,
        1 inside every set_tris, 2 lazily once a scene has served max(32 rays per triangle, 8 Mi
        rays), 3 (default) like 2 plus temporal reuse: set_tris of the same size refits the optimised
        topology."""
        self._need()
        self._check(lib().prt_b200_set_tree_optimisation(self._h, int(mode), int(passes)))

    @property
    def tree_depth(self) -> int:
        """height of the optimised tree; 0 while the current tree is the plain radix tree"""
        self._need()
        return int(lib().prt_b200_tree_depth(self._h))

    @property
    def refits(self) -> int:
        """set_tris calls served by refitting the optimised topology (mode 3)"""
        self._need()
        return int(lib().prt_b200_refits(self._h))

    @property
    def refit_rejects(self) -> int:
        self._need()
        return int(lib().prt_b200_refit_rejects(self._h))

    @property
    def strict_fallbacks(self) -> int:
        self._need()
        return int(lib().prt_b200_strict_fallbacks(self._h))

    @property
    def last_optimise_ms(self) -> float:
        self._need()
        return float(lib().prt_b200_last_optimise_ms(self._h))

    def set_triangle_test(self, mode: int):
        """0 (default): the reference's Moeller-Trumbore arithmetic (core.hpp:27-65), results identical
        to the reference; 1: opt-in watertight test (Woop et al. 2013).  Applies from the next
        set_tris."""
        self._need()
        self._check(lib().prt_b200_set_triangle_test(self._h, int(mode)))

    @property
    def triangle_test(self) -> int:
        self._need()
        return int(lib().prt_b200_triangle_test(self._h))

    def set_wide_nodes(self, mode: int):
        """0 never, 1 always, 2 only for reordered (incoherent) batches; applies from the next set_tris."""
        self._need()
        self._check(lib().prt_b200_set_wide_nodes(self._h, int(mode)))

    def download_wide(self):
        self._need()
        out = np.zeros((self.num_nodes, 64), np.uint8)
        self._check(lib().prt_b200_download_wide(self._h, out.ctypes.data))
        return out

    @property
    def sorted_batches(self):
        return lib().prt_b200_sorted_batches(self._h)

    def read_bandwidth(self, nbytes: int, iters: int = 20) -> float:
        self._need()
        g = C.c_float()
        self._check(lib().prt_b200_read_bandwidth(self._h, int(nbytes), int(iters), C.byref(g)))
        return g.value

    # --- introspection --------------------------------------------------------------------------
    @property
    def num_devices(self) -> int:
        self._need()
        return int(lib().prt_b200_num_devices(self._h))

    @property
    def broadcast_path(self) -> str:
        """what the last multi-GPU set_tris moved the triangles with: "nccl", "p2p" or "" """
        self._need()
        return lib().prt_b200_broadcast_path(self._h).decode()

    @property
    def last_transfer_bytes(self):
        """(host->device, device->host) bytes of the last nearest_hits call"""
        self._need()
        return (int(lib().prt_b200_last_h2d_bytes(self._h)), int(lib().prt_b200_last_d2h_bytes(self._h)))

    @property
    def num_tris(self):
        return lib().prt_b200_num_tris(self._h)

    @property
    def num_nodes(self):
        return lib().prt_b200_num_nodes(self._h)

    @property
    def bvh_root(self):
        return lib().prt_b200_bvh_root(self._h)

    @property
    def bvh_bytes(self):
        return lib().prt_b200_bvh_bytes(self._h)

    @property
    def launch_count(self):
        return lib().prt_b200_launch_count(self._h)

    @property
    def last_build_ms(self):
        return lib().prt_b200_last_build_ms(self._h)

    @property
    def last_trace_ms(self):
        return lib().prt_b200_last_trace_ms(self._h)

    @property
    def last_kernel_ms(self):
        """the traversal kernel of the last trace_dev* call alone (without ray reordering)"""
        return lib().prt_b200_last_kernel_ms(self._h)

    @property
    def exotic_rays(self):
        return lib().prt_b200_exotic_rays(self._h)

    @property
    def graph_replays(self):
        """set_tris calls served by replaying the CUDA graph of the rebuild chain (small scenes)"""
        return lib().prt_b200_graph_replays(self._h)

    @property
    def l2_bytes(self):
        return lib().prt_b200_l2_bytes(self._h)

    def download_bvh(self):
        """-> (nodes (n_nodes,16) float32 view of the 64-byte nodes, tris (n_tris,16) float32: the 64-byte records)."""
        self._need()
        nodes = np.zeros((self.num_nodes, 16), np.float32)
        tris = np.zeros((self.num_tris, 16), np.float32)
        self._check(lib().prt_b200_download_bvh(self._h, nodes.ctypes.data, tris.ctypes.data))
        return nodes, tris

    # --- helpers --------------------------------------------------------------------------------
    def _need(self):
        if not self._h:
            raise RuntimeError("CUDA backend is not initialised: call select_backend(backend) "
                               "or backend.init() first")

    def _check(self, rc):
        if rc:
            raise RuntimeError(f"prt_b200 error {rc}: " +
                               lib().prt_b200_last_error(self._h).decode())

    def __del__(self):
        try:
            self.shutdown()
        except Exception:
            pass


class _Pinned:
    def __init__(self, nbytes):
        self.ptr = lib().prt_b200_alloc_pinned(nbytes)
        if not self.ptr:
            raise MemoryError(f"cudaMallocHost({nbytes}) failed")
        self.buf = (C.c_char * max(1, nbytes)).from_address(self.ptr)

    def __del__(self):
        try:
            lib().prt_b200_free_pinned(self.ptr)
        except Exception:
            pass


def pinned_empty(shape, dtype):
    """numpy array in page-locked host memory (the host entry points then DMA directly)."""
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) * dt.itemsize
    owner = _Pinned(n)
    owner.buf._prt_owner = owner  # arr.base -> ctypes buffer -> owner: freed with the last view
    return np.frombuffer(owner.buf, dtype=np.uint8, count=n).view(dt).reshape(shape)


# --- registry (backend.hpp:39-46, 83-95; src/backend.cpp:50-56) --------------------------------
selected_backend: Backend | None = None
_all_backends: list[Backend] = []
_available_backends: list[Backend] = []


def all_backends():
    return _all_backends


def available_backends():
    return _available_backends


def select_backend(backend: Backend):
    """src/backend.cpp:50-56: shut the previous selection down, then init the new one."""
    global selected_backend
    if selected_backend is not None:
        selected_backend.shutdown()
    selected_backend = backend
    backend.init()


def register_backend(b: Backend):
    """RegisterBackend, backend.hpp:85-95: the first available backend becomes the selection."""
    if b.is_available():
        if selected_backend is None:
            select_backend(b)
        _available_backends.append(b)
    _all_backends.append(b)


def nearest_hits(rays, *tags):
    """Free function, nearest_hits_impl.hpp:28-36: dispatch to the selected backend; raises
    RuntimeError("Unknown backend") when none is selected, like the reference."""
    if selected_backend is None:
        raise RuntimeError("Unknown backend")
    return selected_backend.nearest_hits(rays, *tags)


cuda_backend = CUDABackend()
register_backend(cuda_backend)
