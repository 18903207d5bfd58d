"""Host-pure logic of the concurrency benchmark: grammar, defaults, autotune, verdict, output format.
Runs on a GPU-less box through the deterministic fake backend and the real CPU/OpenMP backend."""
import os
import re
import subprocess

import pytest


def run(native, argv, backend):
    rc, out, err = native.concurency_main(argv, backend)
    return rc, out, err


FAKE = "fake:C=0.01,MD=0.0005,DM=0.0004,HD=0.0002,DH=0.0002,overlap=0.95"


def test_strip_twos(native):
    assert native.strip_twos("M2D") == "MD"
    assert native.strip_twos("D2P") == "DP"
    assert native.strip_twos("C") == "C"


def test_verdict_lines_are_parser_stable(native):
    rc, out, _ = run(native, ["in_order", "--commands", "C", "M2D", "--commands", "M2D", "D2M"], FAKE)
    assert rc == 0
    lines = [l for l in out.splitlines() if l.startswith("## ")]
    assert lines == ["## in_order | C MD | SUCCESS: Close from Theoretical Speedup",
                     "## in_order | MD DM | SUCCESS: Close from Theoretical Speedup"]
    # every line format of the reference's stdout contract is present
    for needle in ["# Performing Autotuning to Balance Commands Times", "Parameters used:",
                   "# in_order | C MD | Starting Benchmarking...", "Minimum Measured Total Time Serial: ",
                   "  Minimum Time Command 0 (  C): ", "  Minimum Time Command 1 ( MD): ",
                   "Maximum Theoretical Speedup: ", "Minimum Measured Total Time //: ",
                   "Speedup Relative to Serial: "]:
        assert needle in out, needle
    assert re.search(r"Minimum Time Command 1 \( MD\): \d+us \([\d.e+]+ GBytes/s\)", out)


def test_failure_when_far_from_theoretical(native):
    spec = "fake:C=0.01,MD=0.0005,overlap=0.0"   # no overlap at all -> speedup 1x vs theoretical 2x
    rc, out, _ = run(native, ["nowait", "--commands", "C", "M2D"], spec)
    assert rc == 1
    assert "## nowait | C MD | FAILURE: Far from Theoretical Speedup" in out


def test_min_bandwidth_floor(native):
    rc, out, _ = run(native, ["nowait", "--min_bandwidth", "1000", "--commands", "H2D", "D2H"], FAKE)
    assert rc == 1
    assert "| FAILURE: Minimun Bandwish not reached" in out
    rc, out, _ = run(native, ["nowait", "--min_bandwidth", "0.001", "--commands", "H2D", "D2H"], FAKE)
    assert rc == 0


def test_autotune_balances_commands_linear_model(native):
    # copies at default 250M elements: MD 125000us, DM 100000us -> target = 100000us (fastest copy)
    rc, out, _ = run(native, ["in_order", "--commands", "C", "M2D", "D2M"], FAKE)
    params = dict(re.findall(r"^\s+(\w+): (\d+)$", out, flags=re.M))
    assert int(params["globalsize_DM"]) == 250_000_000          # already the fastest
    assert int(params["globalsize_MD"]) == 200_000_000          # 250M * 100000/125000
    assert int(params["tripcount_C"]) == 10_000_000             # 40000 * 100000/400
    assert int(params["globalsize_C"]) == 1


def test_user_pinned_parameters_are_not_tuned(native):
    rc, out, _ = run(native, ["in_order", "--tripcount_C", "1234", "--globalsize_M2D", "1000000",
                              "--commands", "C", "M2D"], FAKE)
    assert "# Performing Autotuning" not in out
    assert "tripcount_C: 1234" in out and "globalsize_MD: 1000000" in out


def test_single_unique_command_skips_autotune(native):
    rc, out, _ = run(native, ["in_order", "--commands", "C", "C"], FAKE)
    assert "# Performing Autotuning" not in out
    assert "tripcount_C: 40000" in out


def test_default_memory_size_flag(native):
    rc, out, _ = run(native, ["in_order", "--globalsize_default_memory", "1000", "--tripcount_C", "5",
                              "--globalsize_H2D", "77", "--commands", "H2D", "C", "D2H"], FAKE)
    assert "globalsize_HD: 77" in out and "globalsize_DH: 1000" in out


@pytest.mark.parametrize("argv,msg", [
    ([], "Usage:"),
    (["bogus_mode", "--commands", "C"], "ERROR: Need to specify:"),
    (["in_order"], "ERROR: Need to specify --commands"),
    (["in_order", "--commands", "X2D"], "ERROR: Unsupported value for COMMAND: X2D"),
    (["in_order", "--commands", "H2M"], "ERROR: Unsupported value for COMMAND: H2M"),
    (["in_order", "--commands", "M2H"], "ERROR: Unsupported value for COMMAND: M2H"),
    (["in_order", "--commands", "CC"], "ERROR: Unsupported value for COMMAND: CC"),
    (["in_order", "--frobnicate", "--commands", "C"], "ERROR: Unsupported option: '--frobnicate'"),
    (["in_order", "--queues"], "ERROR: Need to specify a value for '--queues'"),
    (["in_order", "--repetitions", "abc", "--commands", "C"], "ERROR: Invalid integer"),
])
def test_usage_errors_exit_1(native, argv, msg):
    rc, out, _ = run(native, argv, FAKE)
    assert rc == 1 and msg in out and "Usage:" in out


def test_shared_memory_letter_is_accepted(native):
    """The reference documents S (shared) but its CLI rejects it (main.cpp:186); ours accepts it."""
    rc, out, _ = run(native, ["in_order", "--commands", "S2D", "C"], FAKE)
    assert "## in_order | SD C |" in out


def test_unbalance_warning_goes_to_stderr(native):
    rc, out, err = run(native, ["in_order", "--tripcount_C", "1", "--globalsize_M2D", "100000000",
                                "--commands", "C", "M2D"], FAKE)
    assert "WARNING: Large Unbalance Between Commands" in err


def test_judge_function(native):
    assert native.concurency_judge(2.0, 1.9, 10.0, -1.0, 100).startswith("SUCCESS")
    assert native.concurency_judge(2.0, 1.5, 10.0, -1.0, 100).startswith("FAILURE: Far")
    assert native.concurency_judge(2.0, 1.9, 10.0, 50.0, 100).startswith("FAILURE: Minimun")
    assert native.concurency_judge(2.0, 1.9, 10.0, 50.0, 0).startswith("SUCCESS")  # no bytes -> no floor


def test_byte_counter_is_64_bit(native):
    """5 GB aggregate must not wrap (the reference's `unsigned bytes` does, main.cpp:26)."""
    rc, out, _ = run(native, ["in_order", "--globalsize_H2D", "700000000", "--globalsize_D2H", "700000000",
                              "--commands", "H2D", "D2H"], "fake:HD=0.0001,DH=0.0001,overlap=1.0")
    m = re.search(r"Minimum Measured Total Time //: (\d+)us \(([\d.e+]+) GBytes/s\)", out)
    us, bw = int(m.group(1)), float(m.group(2))
    assert abs(bw - 1e-3 * 2 * 700000000 * 4 / us) / bw < 1e-3


def test_json_rows(native, tmp_path):
    import json

    path = tmp_path / "rows.jsonl"
    run(native, ["fused", "--json", str(path), "--commands", "C", "D2P", "--commands", "A", "H2D"], FAKE)
    rows = [json.loads(l) for l in path.read_text().splitlines()]
    assert len(rows) == 2 and rows[0]["commands"] == ["C", "DP"] and rows[0]["mode"] == "fused"
    assert 0.0 <= rows[0]["overlap_fraction"] <= 1.0 and rows[0]["verdict"] in ("SUCCESS", "FAILURE")


def test_cpu_backend_end_to_end(native):
    """BASELINE config #1: compute + copy overlap on the CPU host (OpenMP), no GPU."""
    rc, out, err = run(native, ["host_threads", "--repetitions", "2", "--globalsize_default_memory", "200000",
                                "--tripcount_C", "2000", "--commands", "C", "M2D", "--commands", "M2D", "D2M"],
                       "cpu")
    assert rc in (0, 1)
    assert out.count("## host_threads |") == 2
    rc, out, err = run(native, ["nowait", "--repetitions", "2", "--globalsize_default_memory", "200000",
                                "--commands", "A", "H2D"], "cpu")
    assert "## nowait | A HD |" in out


def test_cpu_bench_api_contract(native):
    r = native.concurency_bench("cpu", "serial", ["C", "MD"], {"globalsize_C": 1, "tripcount_C": 1000,
                                                               "globalsize_MD": 100000}, False, -1, 2, False)
    assert len(r["per_command_us"]) == 2
    assert r["total_us"] <= sum(r["per_command_us"])      # "best theoretical serial" rule
    r = native.concurency_bench("cpu", "host_threads", ["C", "MD"], {"globalsize_C": 1, "tripcount_C": 1000,
                                                                     "globalsize_MD": 100000}, False, -1, 2, False)
    assert r["per_command_us"] == []                       # only filled in serial mode


def test_host_only_binary(bin_dir):
    exe = os.path.join(bin_dir, "omp_con")
    p = subprocess.run([exe, "nowait", "--repetitions", "2", "--globalsize_default_memory", "100000",
                        "--tripcount_C", "500", "--commands", "C", "M2D"], capture_output=True, text=True)
    assert "## nowait | C MD |" in p.stdout
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 1 and "Usage:" in p.stdout


def test_full_binary_falls_back_to_cpu_without_gpu(bin_dir, gpu_count):
    if gpu_count:
        pytest.skip("GPU present")
    exe = os.path.join(bin_dir, "concurency")
    p = subprocess.run([exe, "host_threads", "--repetitions", "1", "--globalsize_default_memory", "100000",
                        "--tripcount_C", "500", "--commands", "C", "M2D"], capture_output=True, text=True)
    assert "using the CPU/OpenMP backend" in p.stderr
    assert "## host_threads | C MD |" in p.stdout
