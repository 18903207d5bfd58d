import json
import os
import subprocess
import sys

import pytest

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
    bad = run("topology", "fake:8:nvswitch")                 # a bad spec is a message and exit 1, not a traceback
    assert bad.returncode == 1 and bad.stderr.startswith("Error:") and "Traceback" not in bad.stderr
    assert "tp " in run("--help").stdout
    log = tmp_path / "l.log"
    log.write_text("export A=1\n## fused | C DP | SUCCESS: Close from Theoretical Speedup\n")
    assert "fused" in run("parse", str(log)).stdout


def test_python_programs_fail_cleanly_without_gpu(gpu_count):
    if gpu_count:
        return
    for prog in ("peer2pear", "allreduce", "halo", "tp", "interop"):
        p = run(prog)
        assert p.returncode == 1 and p.stderr.startswith(f"Error: {prog}: no CUDA device"), p.stderr
        assert "Traceback" not in p.stderr
    assert run("tp", "--help").returncode == 0            # usage is available everywhere


def test_native_clis_fail_cleanly_without_gpu(bin_dir, gpu_count):
    if gpu_count:
        return
    for exe, needle in (("peer2pear", "no CUDA device"), ("allreduce", "No devices"), ("halo", "No devices"),
                        ("interop_torchless", "no CUDA device"), ("interop_driver", "no CUDA device")):
        p = subprocess.run([os.path.join(bin_dir, exe)], capture_output=True, text=True)
        assert p.returncode == 1 and needle in (p.stderr + p.stdout), exe
    p = subprocess.run([os.path.join(bin_dir, "peer2pear"), "--help"], capture_output=True, text=True)
    assert p.returncode == 0 and "--transport" in p.stdout
    p = subprocess.run([os.path.join(bin_dir, "allreduce"), "-h"], capture_output=True, text=True)
    assert p.returncode == 1 and "2^k elements" in p.stdout
    p = subprocess.run([os.path.join(bin_dir, "halo"), "-h"], capture_output=True, text=True)
    assert p.returncode == 1 and "--mode pull|push" in p.stdout
    p = subprocess.run([os.path.join(bin_dir, "halo"), "--mode", "teleport"], capture_output=True, text=True)
    assert p.returncode == 1 and "--mode must be pull or push" in p.stderr
    p = subprocess.run([os.path.join(bin_dir, "peer2pear"), "--transport", "carrier-pigeon"], capture_output=True,
                       text=True)
    assert p.returncode == 1 and "unknown transport" in p.stderr


def test_native_clis_host_only_plumbing(bin_dir):
    """--cpu: the pattern logic (pairing, ring order, verification, output lines) without a GPU."""
    import re

    p = subprocess.run([os.path.join(bin_dir, "peer2pear"), "host", "--cpu", "-n", "4", "--bytes", "65536",
                        "--bytes", "1048576", "--iters", "3"], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    assert len(re.findall(r"host \[\d+ B\] Unidirectional Bandwidth: [\d.e+]+ GB/s", p.stdout)) == 2
    assert len(re.findall(r"host \[\d+ B\] Bidirectional Bandwidth: [\d.e+]+ GB/s", p.stdout)) == 2
    assert "VERIFICATION FAILED" not in p.stdout
    for args, n in ((["-n", "4", "-p", "14"], 4), (["-n", "6", "-p", "10", "-a"], 6),
                    (["-n", "2", "-p", "12", "--type", "int"], 2)):
        p = subprocess.run([os.path.join(bin_dir, "allreduce"), "--cpu", *args], capture_output=True, text=True,
                           timeout=120)
        assert p.returncode == 0, p.stdout + p.stderr
        assert p.stdout.count("Passed") == n and "Elapsed (max over ranks)" in p.stdout
    p = subprocess.run([os.path.join(bin_dir, "allreduce.int"), "--cpu", "-n", "4", "-p", "8"], capture_output=True,
                       text=True, timeout=60)
    assert p.returncode == 0 and " int host-threads" in p.stdout


def test_openmp_sweep_script_end_to_end(bin_dir, tmp_path):
    """scripts/run_omp.sh == the reference's run_omp.sh: env matrix x modes x groups -> log -> tables."""
    root = os.path.dirname(bin_dir)
    p = subprocess.run(["bash", os.path.join(root, "scripts", "run_omp.sh")], cwd=tmp_path, capture_output=True,
                       text=True, timeout=300, env=dict(os.environ, HPCP_OMP_ELEMS="200000"))
    assert p.returncode == 0, p.stderr[-2000:]
    for env in ("OMP_PROC_BIND=false", "OMP_PROC_BIND=spread OMP_PLACES=cores", "OMP_WAIT_POLICY=active"):
        assert env in p.stdout
    assert p.stdout.count("host_threads    nowait") == 3          # one table per environment
    for group in ("C C", "C MD", "C DM", "MD DM", "HD DH"):
        assert p.stdout.count("\n" + group + " ") == 3
    logs = list(tmp_path.glob("tmp-omp-*/omp.log"))
    assert len(logs) == 1 and logs[0].read_text().count("## ") == 30   # 3 envs x 2 modes x 5 groups


def test_p2p_sweep_script_host_only(bin_dir):
    """scripts/p2p_run.sh == the reference's p2p/run.sh (policy x transport x ranks); HPCP_P2P_CPU=1 runs the
    same sweep with host threads so the script itself is testable without a GPU."""
    import re

    root = os.path.dirname(bin_dir)
    p = subprocess.run(["bash", os.path.join(root, "scripts", "p2p_run.sh")], capture_output=True, text=True,
                       timeout=300, env=dict(os.environ, HPCP_P2P_CPU="1", HPCP_NUM_DEVICES="4",
                                             HPCP_P2P_BYTES="65536"))
    assert p.returncode == 0, p.stderr[-2000:]
    lines = re.findall(r"^peer2pear_(\w+) (\d) (\w+) (Uni|Bi)directional Bandwidth: [\d.e+-]+ GB/s$", p.stdout, re.M)
    assert len(lines) == 3 * 4 * 2 * 2                      # policies x transports x rank counts x directions
    assert {l[0] for l in lines} == {"sendrecv", "put", "get", "memcpy"}
    assert {l[2] for l in lines} == {"compact", "spread", "compact_plan"}


def test_cuda_sweep_script_with_cpu_fallback(bin_dir, tmp_path, gpu_count):
    """scripts/run_cuda.sh == the reference's run_sycl.sh.  Without a GPU `bin/concurency` falls back to its CPU
    backend, so the script (env matrix, logging, JSON rows, tables) is exercised here with the two host modes."""
    if gpu_count:
        pytest.skip("GPU present: the script is run on the device by the profiling calls")
    root = os.path.dirname(bin_dir)
    p = subprocess.run(["bash", os.path.join(root, "scripts", "run_cuda.sh")], cwd=tmp_path, capture_output=True,
                       text=True, timeout=600, env=dict(os.environ, HPCP_CUDA_ELEMS="200000",
                                                        HPCP_CUDA_MODES="host_threads nowait"))
    assert p.returncode == 0, p.stderr[-2000:]
    assert p.stdout.count("host_threads    nowait") == 4          # one table per environment of the matrix
    assert "CUDA_DEVICE_MAX_CONNECTIONS=32" in p.stdout and "HPCP_FUSED_COPY_ENGINE=ldst" in p.stdout
    rows = list(tmp_path.glob("tmp-cuda-*/cuda.jsonl"))
    assert len(rows) == 1 and len(rows[0].read_text().splitlines()) == 4 * 2 * 5


def test_reference_program_names_and_json_report_pipeline(bin_dir, tmp_path):
    """The reference's program names exist as links and set the matching defaults; the JSON rows of every
    program feed utils.report (all host-only here)."""
    rows = tmp_path / "rows.jsonl"

    def run(exe, *args):
        p = subprocess.run([os.path.join(bin_dir, exe), *args], capture_output=True, text=True, timeout=120)
        assert p.returncode in (0, 1), p.stdout + p.stderr
        return p.stdout

    for exe, kind, typ in (("allreduce-mpi-sycl.float", "managed", "float"), ("allreduce-mpi-sycl.int", "managed", "int"),
                           ("allreduce-usm-mpi-omp-offload.float", "device", "float"),
                           ("allreduce-map-mpi-omp-offload.float", "mapped-host-malloc", "float")):
        out = run(exe, "--cpu", "-n", "2", "-p", "8", "--json", str(rows))
        assert out.count("Passed") == 2 and f"arrays={kind}" in out and f" {typ} host-threads" in out
    assert "arrays=pinned-host" in run("allreduce-mpi-sycl.float", "--cpu", "-n", "2", "-p", "8", "-H")   # flags still win
    run("peer2pear_i", "two-sided", "--cpu", "-n", "2", "--bytes", "65536", "--json", str(rows))
    run("peer2pear_w", "one-sided", "--cpu", "-n", "2", "--bytes", "65536", "--json", str(rows))
    assert "## nowait | C C |" in run("omp_nowait", "nowait", "--tripcount_C", "1000", "--commands", "C", "C",
                                      "--json", str(rows))
    assert "## host_threads | C C |" in run("omp_host_threads", "host_threads", "--tripcount_C", "1000",
                                            "--commands", "C", "C", "--json", str(rows))
    assert "Usage:" in run("sycl_con")
    parsed = [json.loads(l) for l in rows.read_text().splitlines()]
    assert [r["transport"] for r in parsed if r["pattern"] == "peer2pear"] == ["sendrecv", "put"]
    assert [r["algo"] for r in parsed if r["pattern"] == "allreduce"] == ["host-ring"] * 4
    from hpc_patterns_b200.utils import report

    text = report.render(report.load_rows([str(rows)]))
    assert "## peer2pear" in text and "## allreduce miniapp" in text and "## concurrency bench" in text
    assert "| two-sided | sendrecv | host |" in text and "| one-sided | put | host |" in text


def test_program_options_do_not_collide_with_torchrun_abbreviations():
    """torchrun's argparse resolves ABBREVIATIONS of its own options anywhere on the command line, even after the
    script name: a program option such as `--m` dies with "ambiguous option" (or is silently eaten when it is the
    prefix of exactly one torchrun option).  Every long option of the programs that are launched under torchrun
    must therefore not be a prefix of any torchrun option."""
    import re

    from torch.distributed.run import get_args_parser

    theirs = [o for a in get_args_parser()._actions for o in a.option_strings if o.startswith("--")]
    programs = ["bench.py", "hpc_patterns_b200/models/tensor_parallel.py", "hpc_patterns_b200/models/allreduce.py",
                "hpc_patterns_b200/models/peer2pear.py", "scripts/gemm_put_bench.py"]
    clashes = []
    for rel in programs:
        src = open(os.path.join(ROOT, rel)).read()
        for opt in set(re.findall(r'add_argument\(\s*"(--[A-Za-z0-9_-]+)"', src)):
            hit = [t for t in theirs if t.startswith(opt)]
            if hit:
                clashes.append((rel, opt, hit))
    assert not clashes, clashes
