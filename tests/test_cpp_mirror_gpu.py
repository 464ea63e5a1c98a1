"""The C++ host mirror (include/ndstats_b200.hpp) on a device: its self-test runs the reference's known answers
(tests/sort.rs:35-44, the partition_mut doc-test src/sort.rs:66-88, interpolation at a fraction of 0.5, skipnan, the
min / max family) through the C ABI.  The host-only half runs in the CPU suite (tests/test_cabi_cpu.py)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "ndarray-stats_b200", "cpp")


@pytest.mark.gpu
def test_cpp_mirror_selftest_on_device(gpu_ctx):
    built = subprocess.run(["make", "-C", CPP, "-s", "selftest"], capture_output=True, text=True)
    if built.returncode != 0 or not os.path.exists(os.path.join(CPP, "selftest")):
        pytest.skip("no C++ toolchain on this box: " + built.stderr[-200:])
    r = subprocess.run([os.path.join(CPP, "selftest")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "cpp mirror self-test: ok (with a CUDA device)" in r.stdout
