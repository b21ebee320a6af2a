"""ctypes binding of the C ABI in include/prt_b200.h (portablert_b200/libprt_b200.so).

The shared library is the product; this module fails loudly when it is missing -- there is no
Python or CPU fallback for the hot path.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# PRT_B200_LIB: alternative build of the same library (A/B experiments); default = the in-tree one
LIB_PATH = os.environ.get("PRT_B200_LIB") or os.path.join(HERE, "libprt_b200.so")

OK, E_NO_DEVICE, E_CUDA, E_ARG, E_OOM, E_LIMIT = range(6)


class HitLayout(C.Structure):
    _fields_ = [("stride", C.c_uint32)] + [(n, C.c_int32) for n in (
        "off_u", "off_v", "off_t", "off_pid", "off_valid", "off_px", "off_py", "off_pz")]


class SoaOut(C.Structure):
    _fields_ = [("uv", C.c_void_p), ("t", C.c_void_p), ("pid", C.c_void_p), ("p", C.c_void_p),
                ("valid", C.c_void_p)]


class TraceOpts(C.Structure):
    _fields_ = [("prune", C.c_int), ("slack_rel", C.c_float), ("slack_ulps", C.c_float)]


# every symbol include/prt_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "prt_b200_abi_version": (C.c_int, []),
    "prt_b200_device_count": (C.c_int, []),
    "prt_b200_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "prt_b200_create_multi": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "prt_b200_num_devices": (C.c_int, [C.c_void_p]),
    "prt_b200_broadcast_path": (C.c_char_p, [C.c_void_p]),
    "prt_b200_last_h2d_bytes": (C.c_uint64, [C.c_void_p]),
    "prt_b200_last_d2h_bytes": (C.c_uint64, [C.c_void_p]),
    "prt_b200_destroy": (None, [C.c_void_p]),
    "prt_b200_device_name": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t]),
    "prt_b200_set_tris": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "prt_b200_set_tris_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_float)]),
    "prt_b200_nearest_hits": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32,
                                        C.POINTER(HitLayout), C.c_void_p]),
    "prt_b200_trace_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32,
                                     C.POINTER(SoaOut), C.POINTER(C.c_float)]),
    "prt_b200_trace_dev_aos": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32,
                                         C.POINTER(HitLayout), C.c_void_p, C.POINTER(C.c_float)]),
    "prt_b200_set_trace_opts": (C.c_int, [C.c_void_p, C.POINTER(TraceOpts)]),
    "prt_b200_set_tree_optimisation": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "prt_b200_tree_depth": (C.c_int32, [C.c_void_p]),
    "prt_b200_last_optimise_ms": (C.c_float, [C.c_void_p]),
    "prt_b200_strict_fallbacks": (C.c_uint64, [C.c_void_p]),
    "prt_b200_refits": (C.c_uint64, [C.c_void_p]),
    "prt_b200_refit_rejects": (C.c_uint64, [C.c_void_p]),
    "prt_b200_set_triangle_test": (C.c_int, [C.c_void_p, C.c_int]),
    "prt_b200_triangle_test": (C.c_int, [C.c_void_p]),
    "prt_b200_set_ray_sorting": (C.c_int, [C.c_void_p, C.c_int]),
    "prt_b200_set_wide_nodes": (C.c_int, [C.c_void_p, C.c_int]),
    "prt_b200_download_wide": (C.c_int, [C.c_void_p, C.c_void_p]),
    "prt_b200_sorted_batches": (C.c_uint64, [C.c_void_p]),
    "prt_b200_trace_count_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "prt_b200_num_tris": (C.c_uint64, [C.c_void_p]),
    "prt_b200_num_nodes": (C.c_uint64, [C.c_void_p]),
    "prt_b200_bvh_root": (C.c_int32, [C.c_void_p]),
    "prt_b200_bvh_bytes": (C.c_uint64, [C.c_void_p]),
    "prt_b200_launch_count": (C.c_uint64, [C.c_void_p]),
    "prt_b200_last_build_ms": (C.c_float, [C.c_void_p]),
    "prt_b200_last_trace_ms": (C.c_float, [C.c_void_p]),
    "prt_b200_last_kernel_ms": (C.c_float, [C.c_void_p]),
    "prt_b200_exotic_rays": (C.c_uint64, [C.c_void_p]),
    "prt_b200_l2_bytes": (C.c_uint64, [C.c_void_p]),
    "prt_b200_graph_replays": (C.c_uint64, [C.c_void_p]),
    "prt_b200_download_bvh": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "prt_b200_read_bandwidth": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_float)]),
    "prt_b200_alloc_pinned": (C.c_void_p, [C.c_size_t]),
    "prt_b200_free_pinned": (None, [C.c_void_p]),
    "prt_b200_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "prt_b200_host_unregister": (C.c_int, [C.c_void_p]),
    "prt_b200_last_error": (C.c_char_p, [C.c_void_p]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C portablert_b200/csrc` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
                "portablert_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            if os.environ.get("PRT_B200_LIB_OLD_ABI") and not hasattr(L, name):
                continue  # A/B runs against a library built from an older commit (tools/sweep.py)
            fn = getattr(L, name)  # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
