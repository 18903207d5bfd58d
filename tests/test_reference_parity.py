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
        line = re.sub(r"-?nan|\binf", "N", line)       # a 0 us measurement divides by zero in both programs
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


# ------------------------------------------------------------------------------------------------ differential fuzz ----
from hypothesis import HealthCheck, given, settings, strategies as st  # noqa: E402

# Not in the alphabet on purpose: two-letter tokens with a `C` ("CC", "CM", ...) and the empty token.  The reference
# accepts them (its check is "letters from CMDH, not HM / MH", main.cpp:186-191) and then times a meaningless copy
# "from C to C" with globalsize_CC = 1, or a nameless command; here they are usage errors.
_TOKENS = ["C", "M2D", "D2M", "MD", "DM", "H2D", "D2H", "HD", "DH", "C2", "2C", "M2D2", "HM", "MH", "X", "M2X"]
_groups = st.lists(st.lists(st.sampled_from(_TOKENS), min_size=0, max_size=3), min_size=0, max_size=3)
_flags = st.lists(st.sampled_from([["--verbose"], ["--enable_profiling"], ["--queues", "1"], ["--queues", "2"],
                                   ["--repetitions", "1"], ["--min_bandwidth", "0.000001"], ["--bogus"], ["-x"],
                                   ["--tripcount_C", "7"], ["--globalsize_C", "2"]]),
                  min_size=0, max_size=3)


@given(groups=_groups, flags=_flags, mode=st.sampled_from(["nowait", "host_threads", "serial", "in_order", ""]))
@settings(max_examples=60, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture])
def test_random_command_lines_behave_like_the_reference(ref_bins, bin_dir, groups, flags, mode):
    """Random argv (valid and invalid) through the reference binary and ours: same exit status, same usage-or-run
    decision, same sequence of line shapes.  Sizes are tiny, the verdict text itself is normalised away.
    (derandomize: the suite runs the same 60 command lines every time; drop the flag locally to explore other seeds; 8 x 60 more were run when this was written.)"""
    argv = [mode] if mode else []
    for f in flags:
        argv += f
    argv += ["--globalsize_default_memory", "2000", "--tripcount_C", "50", "--repetitions", "2"]
    for cmd in ("MD", "DM", "HD", "DH"):       # explicit sizes: no autotuning, whose result depends on 0 us timings
        argv += ["--globalsize_" + cmd, "2000"]
    for g in groups:
        argv += ["--commands"] + list(g)
    ref = subprocess.run([ref_bins.get(mode, ref_bins["nowait"])] + argv, capture_output=True, text=True, timeout=120)
    ours = subprocess.run([os.path.join(bin_dir, "omp_con")] + argv, capture_output=True, text=True, timeout=120)
    ref_usage = "--commands" in ref.stdout and "## " not in ref.stdout and ref.returncode == 1
    ours_usage = "--commands" in ours.stdout and "## " not in ours.stdout and ours.returncode == 1
    assert ours.returncode in (0, 1), (argv, ours.returncode, ours.stderr[-300:])                 # we never crash
    if ref.returncode < 0:
        # the reference's offload allocator aborts ("Wrong Allocation") for device buffers when g++ has no offload
        # device to fall back to — an artefact of building it for the host; nothing to compare with
        return
    assert ref_usage == ours_usage, (argv, ref.stdout[-300:], ours.stdout[-300:])
    if not ref_usage:
        assert (ref.returncode in (0, 1)) and (ours.returncode in (0, 1)), (argv, ref.returncode, ours.returncode)
        assert _shape(ours.stdout) == _shape(ref.stdout), (argv, ref.stdout, ours.stdout)


_modes = st.sampled_from(["nowait", "host_threads", "in_order", "out_of_order", "fused", "serial"])
_cmds = st.lists(st.sampled_from(["C", "MD", "DM", "HD", "DH", "DP", "A", "T"]), min_size=1, max_size=3).map(" ".join)
_verdict = st.tuples(_modes, _cmds, st.sampled_from(["SUCCESS: Close from Theoretical Speedup",
                                                     "FAILURE: Far from Theoretical Speedup",
                                                     "FAILURE: Minimun Bandwish not reached"])
                     ).map(lambda t: f"## {t[0]} | {t[1]} | {t[2]}")
_export = st.lists(st.sampled_from(["CUDA_VISIBLE_DEVICES=0", "HPCP_FUSED_COPY_ENGINE=1", "OMP_PROC_BIND=false",
                                    "CUDA_DEVICE_MAX_CONNECTIONS=32", "A=1 B=2"]), min_size=1, max_size=2
                   ).map(lambda v: "+ export " + " ".join(v))
_noise = st.sampled_from(["# nowait | C MD | Starting Benchmarking...", "Minimum Measured Total Time Serial: 12us",
                          "  Minimum Time Command 0 (  C): 7us", "Speedup Relative to Serial: 1.9x", "",
                          "Parameters used:", "  tripcount_C: 40000", "+ ./omp_con nowait --commands C M2D"])


@given(lines=st.lists(st.one_of(_verdict, _verdict, _export, _noise), min_size=0, max_size=25),
       fmt=st.sampled_from(["simple", "github", "plain"]))
@settings(max_examples=40, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture])
def test_random_logs_render_like_the_reference_parser(tmp_path, lines, fmt):
    """The reference's parse.py and ours print the same tables for any log a sweep script can produce."""
    log = tmp_path / "fuzz.log"
    log.write_text("\n".join(lines) + "\n")
    theirs = subprocess.run([sys.executable, os.path.join(REF_CON, "parse.py"), str(log), fmt], capture_output=True,
                            text=True, timeout=60)
    mine = subprocess.run([sys.executable, "-m", "hpc_patterns_b200.utils.parse", str(log), fmt], capture_output=True,
                          text=True, env=dict(os.environ, PYTHONPATH=ROOT), timeout=60)
    assert theirs.returncode == 0 and mine.returncode == 0, theirs.stderr + mine.stderr
    assert mine.stdout == theirs.stdout, (lines, theirs.stdout, mine.stdout)
