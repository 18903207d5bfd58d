import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu


def test_torch_native_interop_direct(native, capsys):
    from hpc_patterns_b200.models import interop

    interop.demo_direct()
    out = capsys.readouterr().out
    assert "Torch -> HPCP" in out and "HPCP -> Torch" in out and out.strip().endswith("Computation Done")


def test_native_handles_table(native):
    from hpc_patterns_b200.models import interop

    info = interop.demo_native_handles(verbose=False)
    assert info["cu_context"] != 0
    assert interop.get_infos_devices() is interop.get_infos_devices()   # cached table


@pytest.mark.parametrize("exe", ["interop_torchless", "interop_driver"])
def test_native_interop_binaries(bin_dir, exe):
    p = subprocess.run([os.path.join(bin_dir, exe)], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    assert p.stdout.strip().endswith("Computation Done")
