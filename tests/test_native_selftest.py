import os
import subprocess

import pytest


def test_native_selftest_binary(bin_dir):
    """C++ unit tests of the thread-per-rank runtime, device subsets, dtype traits, topology and driver helpers."""
    p = subprocess.run([os.path.join(bin_dir, "native_selftest")], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "native selftest: OK" in p.stdout


def test_host_programs_are_clean_under_asan_and_ubsan(tmp_path):
    """scripts/sanitize_host.sh: AddressSanitizer + UBSan over the host-only concurrency bench and the self-test."""
    import shutil
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    probe = subprocess.run(["/usr/bin/g++", "-fsanitize=address,undefined", "-x", "c++", "-", "-o",
                            str(tmp_path / "probe")], input="int main(){}", text=True, capture_output=True)
    if probe.returncode != 0 or shutil.which("bash") is None:
        pytest.skip("no sanitizer runtime in this image")
    p = subprocess.run(["bash", os.path.join(root, "scripts", "sanitize_host.sh"), str(tmp_path / "san")],
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "sanitize_host: OK" in p.stdout, p.stdout + p.stderr
