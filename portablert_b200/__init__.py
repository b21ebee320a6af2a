"""portablert_b200 -- B200-native (sm_100a) CUDA backend for portableRT's nearest-hit path.

The product is ``libprt_b200.so`` (hand-written CUDA kernels behind the C ABI of
``include/prt_b200.h``).  This package is the thin Python host mirror of the reference's backend
interface (see backend.py) used by the tests and by bench.py; it fails loudly if the shared
library has not been built and never falls back to a CPU implementation.

``hitreg`` (record layouts) and ``scenes`` (synthetic inputs) are pure-python helpers and import
without the shared library -- the CPU reference arm of bench.py uses them and must not map the
product.  Everything else (``CUDABackend``, ``select_backend``, ``nearest_hits``, ``lib`` ...) loads
``libprt_b200.so`` on first access and raises ImportError when it is missing.
"""
from . import hitreg, scenes  # noqa: F401  (pure-python helpers: no shared library involved)

_BACKEND_NAMES = ("Backend", "CUDABackend", "all_backends", "available_backends", "cuda_backend",
                  "nearest_hits", "register_backend", "select_backend", "pinned_empty")


def _sub(name):
    import importlib
    return importlib.import_module("." + name, __name__)


def __getattr__(name):
    if name in ("lib", "LIB_PATH", "_lib"):
        _lib = _sub("_lib")
        if name == "lib":
            _lib.lib()  # a missing/broken extension is an import error, not a silent fallback
        return _lib if name == "_lib" else getattr(_lib, name)
    if name in _BACKEND_NAMES or name == "backend":
        _sub("_lib").lib()
        backend = _sub("backend")
        return backend if name == "backend" else getattr(backend, name)
    if name == "sharding":
        return _sub("sharding")
    raise AttributeError(f"module 'portablert_b200' has no attribute {name!r}")


def selected_backend():
    return _sub("backend").selected_backend
