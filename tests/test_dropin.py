"""The reference's own C++ API with this backend plugged in (tests/dropin/dropin_test.cpp):
select_backend / available_backends / set_tris / nearest_hits<Tags...> (free and member form),
CPU backend and CUDA backend side by side in one process."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "dropin_test")


def _run():
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/dropin_test not built (needs /root/reference: make -C tests/dropin)")
    return subprocess.run([BIN], capture_output=True, text=True, timeout=600)


def test_plugin_registers_next_to_the_cpu_backend():
    """Runs everywhere: the backend is compiled in, is listed by all_backends(), reports
    availability truthfully during static initialisation and leaves CPU the default selection."""
    out = _run()
    assert out.returncode == 0, out.stdout + out.stderr
    assert "compiled backends: 2" in out.stdout and out.stdout.strip().endswith("PASS") or \
        "PASS" in out.stdout.splitlines()[-1]


@pytest.mark.gpu
def test_reference_api_cuda_vs_cpu_backend_all_31_combos():
    out = _run()
    assert out.returncode == 0, out.stdout + out.stderr
    assert "CUDA available: 1" in out.stdout, out.stdout
    summary = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert summary["failures"] == 0 and summary["rays_compared"] > 31 * 60000
    assert summary["negative_t"] > 100
    print(out.stdout)


ACCEPT = os.path.join(ROOT, "oracle", "_ref", "acceptance")


def _acceptance(tmp_path, with_bunny=True):
    """Runs the acceptance program (tests/dropin/acceptance.cpp) on the bunny of tests/golden/."""
    import numpy as np
    if not os.path.exists(ACCEPT):
        pytest.skip("oracle/_ref/acceptance not built (needs /root/reference: make -C tests/dropin)")
    args = [ACCEPT, "--csv", str(tmp_path / "results.csv")]
    if with_bunny:
        g = os.path.join(ROOT, "tests", "golden")
        tris = np.load(os.path.join(g, "bunny.npz"))["tris"].astype(np.float32)
        mask = np.unpackbits(np.load(os.path.join(g, "bunny_mask.npz"))["mask"])[:1024 * 1024]
        tris.tofile(tmp_path / "tris.bin")
        (mask * 255).astype(np.uint8).tofile(tmp_path / "mask.bin")
        args += ["--bunny", str(tmp_path / "tris.bin"), str(tmp_path / "mask.bin")]
    out = subprocess.run(args, capture_output=True, text=True, timeout=900)
    rows = [l.split(",") for l in open(tmp_path / "results.csv").read().splitlines()]
    return out, rows


def test_acceptance_program_writes_the_reference_results_csv(tmp_path):
    """examples/validation/main.cpp:249-261: header, one row per check, last column 1 = passed; here
    without a GPU for the reference's own CPU backend (known answers + 1024^2 bunny mask + the
    1 s build / traverse gates), which also pins the program itself to the reference."""
    out, rows = _acceptance(tmp_path)
    assert out.returncode == 0, out.stdout + out.stderr
    assert rows[0] == ["Backend", "Device", "Test", "Subtest", "Value", "Expected", "Validation"]
    assert "Printing all compiled backends:" in out.stdout and "\tCUDA" in out.stdout
    assert "Embree CPU backend:" in out.stdout  # wired in when Embree 4 exists, reported otherwise
    cpu = [r for r in rows[1:] if r[2] == "CPU"]
    assert len(cpu) == 14 and all(r[-1] == "1" for r in cpu), cpu
    assert {r[1] for r in cpu} >= {"t", "u", "v", "primitive_id", "Pixel Validation", "BVH Build time"}
    assert "rays/s" in out.stdout


@pytest.mark.gpu
def test_acceptance_program_on_the_cuda_backend(tmp_path):
    out, rows = _acceptance(tmp_path)
    assert out.returncode == 0, out.stdout + out.stderr
    cuda = [r for r in rows[1:] if r[2] == "CUDA"]
    assert len(cuda) == 14 and all(r[-1] == "1" for r in cuda), cuda
    assert "Testing CUDA" in out.stdout and "B200" in out.stdout
    print(out.stdout)


REF = "/root/reference"


def test_use_cuda_block_configures_and_builds_in_the_reference_cmake(tmp_path):
    """The USE_CUDA block (cmake/portableRT_use_cuda.cmake) inside the reference's OWN
    CMakeLists.txt: a staged copy of the checkout (patched by tools/patch_reference.py; staged under
    tmp, never committed) with the block included where the other backend blocks sit, configured
    with -DUSE_CUDA=ON and built.  The examples sub-directory is left out: it needs SDL2 and four
    network fetches.  CPU only: nothing runs, the plugin just has to compile and link in."""
    import shutil
    import sys
    lib = os.path.join(ROOT, "portablert_b200", "libprt_b200.so")
    cmake = shutil.which("cmake")
    if not os.path.isdir(REF) or cmake is None or not os.path.exists(lib):
        pytest.skip("needs /root/reference, cmake and the built library")
    stage = tmp_path / "portableRT"
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "patch_reference.py"), REF, str(stage)],
                   check=True, capture_output=True)
    shutil.copy(os.path.join(REF, "src", "portableRT.cpp"), stage / "src")
    shutil.copy(os.path.join(REF, "version.hpp.in"), stage)
    text = open(os.path.join(REF, "CMakeLists.txt")).read()
    assert "add_subdirectory(examples)" in text
    text = text.replace("add_subdirectory(examples)", "")
    marker = "set_target_properties(portableRT PROPERTIES \n    RUNTIME_OUTPUT_DIRECTORY"
    assert marker in text, "the reference's CMakeLists.txt changed shape"
    block = os.path.join(ROOT, "cmake", "portableRT_use_cuda.cmake")
    text = text.replace(marker, f"include({block})\n\n" + marker)
    (stage / "CMakeLists.txt").write_text(text)
    build = tmp_path / "build"
    cfg = subprocess.run([cmake, "-S", str(stage), "-B", str(build), "-DUSE_CUDA=ON",
                          f"-DPRT_B200_ROOT={ROOT}", "-DCMAKE_BUILD_TYPE=Release",
                          "-DCMAKE_CXX_FLAGS=-include cstdint"],
                         capture_output=True, text=True, timeout=300)
    assert cfg.returncode == 0, cfg.stdout + cfg.stderr
    cache = (build / "CMakeCache.txt").read_text()
    assert "libprt_b200.so" in cache and "USE_CUDA:BOOL=ON" in cache
    bld = subprocess.run([cmake, "--build", str(build), "-j", "4"], capture_output=True, text=True,
                         timeout=600)
    assert bld.returncode == 0, bld.stdout[-3000:] + bld.stderr[-3000:]
    archive = build / "lib" / "libportableRT.a"
    assert archive.exists()
    syms = subprocess.run(["nm", "-C", str(archive)], capture_output=True, text=True).stdout
    assert "CUDABackend" in syms and "prt_b200_create" in syms  # the plugin is in, the C ABI referenced
