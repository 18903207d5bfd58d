import os
import subprocess


def test_native_selftest_binary(bin_dir):
    """C++ unit tests of the thread-per-rank runtime, device subsets, dtype traits, topology and driver helpers."""
    p = subprocess.run([os.path.join(bin_dir, "native_selftest")], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "native selftest: OK" in p.stdout
