"""The C-ABI library loads on a CPU-only box and exports every symbol include/prt_b200.h declares;
without a GPU it refuses to work instead of falling back to anything."""
import ctypes as C
import os
import re

import pytest

import portablert_b200 as prt
from portablert_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "prt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(prt_b200_\w+)\s*\(", src)))


def test_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 18
    raw = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"libprt_b200.so does not export {n}"
    assert sorted(_lib.SYMBOLS) == names, "python binding table out of sync with the header"
    assert prt.lib().prt_b200_abi_version() == 3


def test_library_contains_sm100a_code_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_no_gpu_means_no_backend_not_a_fallback():
    if prt.lib().prt_b200_device_count() > 0:
        pytest.skip("a B200 is present")
    assert prt.available_backends() == [] and prt.all_backends() == [prt.cuda_backend]
    h = C.c_void_p()
    assert prt.lib().prt_b200_create(C.byref(h), -1) == _lib.E_NO_DEVICE
    assert b"no CPU fallback" in prt.lib().prt_b200_last_error(None)
    with pytest.raises(RuntimeError, match="Unknown backend"):  # nearest_hits_impl.hpp:31-32
        prt.nearest_hits([[0, 0, 0, 0, 0, 1]])
    with pytest.raises(RuntimeError):
        prt.cuda_backend.init()
    with pytest.raises(RuntimeError, match="not initialised"):
        prt.cuda_backend.set_tris([[0] * 9])


def test_null_safety():
    L = prt.lib()
    L.prt_b200_destroy(None)
    assert L.prt_b200_num_tris(None) == 0 and L.prt_b200_launch_count(None) == 0
    assert L.prt_b200_create(None, -1) == _lib.E_ARG


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "portablert_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "liboracle" not in txt and "libprt_ref" not in txt and "libprt_emu" not in txt, f


def test_cmake_build_lists_the_same_sources_as_the_makefile():
    """CMakeLists.txt (for CMake checkouts like the reference) must not drift from the in-tree
    Makefile the tests and the bench use."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    mk = open(os.path.join(root, "portablert_b200", "csrc", "Makefile")).read()
    cm = open(os.path.join(root, "CMakeLists.txt")).read()
    mk_src = set(re.findall(r"\$\(HERE\)(\w+\.cu)", mk))
    cm_src = set(re.findall(r"\$\{SRC\}/(\w+\.cu)", cm))
    assert mk_src == cm_src and "100a" in cm and "USE_CUDA" in open(
        os.path.join(root, "cmake", "portableRT_use_cuda.cmake")).read()
