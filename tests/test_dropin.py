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
