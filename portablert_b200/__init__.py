"""portablert_b200 -- B200-native (sm_100a) CUDA backend for portableRT's nearest-hit path.

The product is ``libprt_b200.so`` (hand-written CUDA kernels behind the C ABI of
``include/prt_b200.h``).  This package is the thin Python host mirror of the reference's backend
interface (see backend.py) used by the tests and by bench.py; it fails loudly if the shared
library has not been built and never falls back to a CPU implementation.
"""
from . import hitreg, scenes  # noqa: F401  (pure-python helpers)
from ._lib import LIB_PATH, lib  # noqa: F401

lib()  # load now: a missing/broken extension must be an import error, not a silent fallback

from .backend import (Backend, CUDABackend, all_backends, available_backends,  # noqa: E402,F401
                      cuda_backend, nearest_hits, register_backend, select_backend)
from . import backend as _backend  # noqa: E402


def selected_backend():
    return _backend.selected_backend
