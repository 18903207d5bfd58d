"""Drop-in parity against the REAL reference program, where it can be built here.

The reference's OpenMP backend (concurency/bench_omp.cpp + main.cpp) compiles with plain ``g++ -fopenmp``
(target regions fall back to the host) once the Intel-only ``omp_target_alloc_host`` is mapped to
``omp_target_alloc`` on the command line — the sources are used unmodified, straight from the read-only
mount.  That gives a ground truth for the stdout contract:

* our binary and the reference binary print the same sequence of line shapes for the same arguments;
* the reference's own ``parse.py`` turns OUR log into the same table our parser renders, and vice versa.

Skipped when the reference tree is not mounted (e.g. on the GPU box).
"""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("HPCP_REFERENCE", "/root/reference")
REF_CON = os.path.join(REF, "concurency")

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF_CON, "bench_omp.cpp")),
                                reason="reference tree not mounted")


@pytest.fixture(scope="module")
def ref_bins(tmp_path_factory):
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    out = tmp_path_factory.mktemp("refomp")
    bins = {}
    for mode, flag in (("host_threads", "HOST_THREADS"), ("nowait", "NOWAIT")):
        exe = out / f"omp_{mode}"
        p = subprocess.run([cxx, "-O2", "-std=c++17", "-fopenmp", f"-D{flag}",
                            "-Domp_target_alloc_host=omp_target_alloc",      # Intel extension -> standard call
                            os.path.join(REF_CON, "main.cpp"), os.path.join(REF_CON, "bench_omp.cpp"), "-o", str(exe)],
                           capture_output=True, text=True, timeout=600)
        if p.returncode != 0:
            pytest.skip("the reference's OpenMP bench does not build with this g++: " + p.stderr[-300:])
        bins[mode] = str(exe)
    return bins


def _shape(text: str):
    """Line shapes: numbers -> N, verdict text kept up to the colon (the result itself is timing-dependent)."""
    out = []
    for line in text.splitlines():
        if line.startswith("# CUDA backend unavailable"):     # our one informational extra line ('#' = comment)
            continue
        if "WARNING: Large Unbalance" in line:                 # timing-dependent in both programs
            continue
        line = re.sub(r"\d+(\.\d+)?(e[+-]?\d+)?", "N", line)
        line = re.sub(r"(SUCCESS|FAILURE): .*", "VERDICT", line)
        out.append(line)
    return out


ARGS = ["--tripcount_C", "3000", "--commands", "C", "C"]


@pytest.mark.parametrize("extra", [[], ["--verbose", "--repetitions", "3"], ["--min_bandwidth", "100000"],
                                   ["--queues", "1", "--verbose", "--repetitions", "2"]])
@pytest.mark.parametrize("mode", ["host_threads", "nowait"])
def test_same_stdout_shape_as_the_reference_binary(ref_bins, bin_dir, mode, extra):
    ref = subprocess.run([ref_bins[mode], mode] + extra + ARGS, capture_output=True, text=True, timeout=300)
    ours = subprocess.run([os.path.join(bin_dir, "omp_con"), mode] + extra + ARGS, capture_output=True, text=True,
                          timeout=300)
    assert "## " + mode + " | C C | " in ref.stdout, ref.stdout + ref.stderr
    assert _shape(ours.stdout) == _shape(ref.stdout)
    assert ref.returncode in (0, 1) and ours.returncode in (0, 1)      # 1 = some group FAILED (same rule)


def test_usage_and_exit_status_match(ref_bins, bin_dir):
    ref = subprocess.run([ref_bins["nowait"]], capture_output=True, text=True)
    ours = subprocess.run([os.path.join(bin_dir, "omp_con")], capture_output=True, text=True)
    assert ref.returncode == ours.returncode == 1
    for flag in ("--commands", "--repetitions", "--min_bandwidth", "--queues", "--enable_profiling"):
        assert flag in ref.stdout and flag in ours.stdout
    for bad in (["bogus_mode", "--commands", "C", "C"], ["nowait", "--bogus"], ["nowait", "--commands", "C", "X"]):
        r = subprocess.run([ref_bins["nowait"]] + bad, capture_output=True, text=True)
        o = subprocess.run([os.path.join(bin_dir, "omp_con")] + bad, capture_output=True, text=True)
        assert r.returncode == o.returncode == 1
        assert r.stdout.splitlines()[0] == o.stdout.splitlines()[0]          # same ERROR line, then the usage text
        assert r.stdout.splitlines()[0].startswith("ERROR: ")
    bad_ref = subprocess.run([ref_bins["nowait"], "nowait", "--commands", "C", "H2M"], capture_output=True, text=True)
    bad_ours = subprocess.run([os.path.join(bin_dir, "omp_con"), "nowait", "--commands", "C", "H2M"],
                              capture_output=True, text=True)
    assert bad_ref.returncode == bad_ours.returncode == 1              # HM / MH are rejected by both


def test_parsers_are_interchangeable(ref_bins, bin_dir, tmp_path):
    logs = {}
    for who, exe in (("ref", None), ("ours", os.path.join(bin_dir, "omp_con"))):
        text = "+ export OMP_PROC_BIND=false\n"
        for mode in ("nowait", "host_threads"):
            p = subprocess.run([exe or ref_bins[mode], mode] + ARGS, capture_output=True, text=True, timeout=300)
            text += p.stdout
        path = tmp_path / f"{who}.log"
        path.write_text(text)
        logs[who] = str(path)
    env = dict(os.environ, PYTHONPATH=ROOT)
    for log in logs.values():
        theirs = subprocess.run([sys.executable, os.path.join(REF_CON, "parse.py"), log], capture_output=True, text=True)
        mine = subprocess.run([sys.executable, "-m", "hpc_patterns_b200.utils.parse", log], capture_output=True,
                              text=True, env=env)
        assert theirs.returncode == 0 and mine.returncode == 0, theirs.stderr + mine.stderr
        assert mine.stdout == theirs.stdout and "OMP_PROC_BIND=false" in mine.stdout
