import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args, env=None):
    e = dict(os.environ, PYTHONPATH=ROOT)
    e.update(env or {})
    return subprocess.run([sys.executable, "-m", "hpc_patterns_b200", *args], capture_output=True, text=True,
                          env=e, cwd=ROOT, timeout=300)


def test_help_and_unknown():
    assert "concurency" in run("--help").stdout
    p = run("nonsense")
    assert p.returncode == 2 and "unknown program" in p.stderr


def test_concurency_through_front_door(native):
    p = run("concurency", "nowait", "--backend", "fake:C=0.01,MD=0.0005,overlap=0.9", "--commands", "C", "M2D")
    assert p.returncode == 0 and "## nowait | C MD | SUCCESS" in p.stdout


def test_topology_and_parse(native, tmp_path):
    p = run("topology", "4:switch")
    assert json.loads(p.stdout)["planes"] == [[0, 1, 2, 3]]
    log = tmp_path / "l.log"
    log.write_text("export A=1\n## fused | C DP | SUCCESS: Close from Theoretical Speedup\n")
    assert "fused" in run("parse", str(log)).stdout


def test_native_clis_fail_cleanly_without_gpu(bin_dir, gpu_count):
    if gpu_count:
        return
    for exe, needle in (("peer2pear", "no CUDA device"), ("allreduce", "No devices"),
                        ("interop_torchless", "no CUDA device"), ("interop_driver", "no CUDA device")):
        p = subprocess.run([os.path.join(bin_dir, exe)], capture_output=True, text=True)
        assert p.returncode == 1 and needle in (p.stderr + p.stdout), exe
    p = subprocess.run([os.path.join(bin_dir, "peer2pear"), "--help"], capture_output=True, text=True)
    assert p.returncode == 0 and "--transport" in p.stdout
    p = subprocess.run([os.path.join(bin_dir, "allreduce"), "-h"], capture_output=True, text=True)
    assert p.returncode == 1 and "2^k elements" in p.stdout
    p = subprocess.run([os.path.join(bin_dir, "peer2pear"), "--transport", "carrier-pigeon"], capture_output=True,
                       text=True)
    assert p.returncode == 1 and "unknown transport" in p.stderr
