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
