"""`__graft_entry__.build()` must rebuild everything from a tree with no artefacts at all.

The GPU boxes receive prebuilt `build/`, `bin/` and `_C*.so` with the snapshot (they are git-ignored, not
gpurun-ignored), so nothing there ever proves that a clean checkout still builds.  This test copies the SOURCE tree
(no build/, bin/, *.so, *.o) to a scratch directory, runs build() there in a fresh interpreter and uses what it built.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_build_from_a_tree_without_artefacts(tmp_path):
    dst = tmp_path / "tree"
    skip_dirs = {"build", "bin", "gpurun_out", ".git", "__pycache__", ".pytest_cache", ".hypothesis", "profiles", "_ref",
                 "sass"}

    def ignore(directory, names):
        out = [n for n in names if n in skip_dirs or n.endswith((".so", ".o", ".a", ".ncu-rep"))]
        return out

    shutil.copytree(ROOT, dst, ignore=ignore)
    assert not (dst / "build").exists() and not (dst / "bin").exists()
    assert not list((dst / "hpc_patterns_b200").glob("_C*.so"))
    env = dict(os.environ)
    env.pop("PYTHONPATH", None)
    p = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.build(); "
                        "import hpc_patterns_b200 as h; C = h.native(); "
                        "print('built', C.__file__, C.strip_twos('H2D'), C.HALO_FLAG_BYTES)"],
                       cwd=dst, env=env, capture_output=True, text=True, timeout=1500)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("built")][-1].split()
    assert line[1].startswith(str(dst)), "the freshly built extension must be the one that was imported"
    assert line[2] == "HD"
    for exe in ("concurency", "omp_con", "peer2pear", "topology", "allreduce", "halo", "interop_torchless",
                "interop_driver", "native_selftest", "allreduce.double", "sycl_con", "peer2pear_w"):
        assert (dst / "bin" / exe).exists(), exe
    # the host-only program of the clean build runs
    r = subprocess.run([str(dst / "bin" / "omp_con"), "nowait", "--globalsize_default_memory", "100000", "--commands", "C",
                        "M2D"], capture_output=True, text=True, timeout=120)
    assert "## nowait | C MD |" in r.stdout, r.stdout + r.stderr
    r = subprocess.run([str(dst / "bin" / "halo"), "-h"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "--mode pull|push" in r.stdout
