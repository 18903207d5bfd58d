"""Property-based tests (hypothesis) of the host-pure logic: the C++ and Python twins must agree and keep
their invariants for arbitrary inputs, not only for the hand-picked cases of the other test files."""
import pytest

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st   # noqa: E402

from hpc_patterns_b200.parallel import tile_mapping as tm   # noqa: E402

FAST = settings(max_examples=200, deadline=None)


@FAST
@given(st.text(alphabet="CMDHSPAT2", max_size=8))
def test_strip_twos_removes_exactly_the_twos(native, token):
    out = native.strip_twos(token)
    assert out == token.replace("2", "") and native.strip_twos(out) == out


@FAST
@given(st.sampled_from(["compact", "spread"]), st.integers(1, 16), st.integers(0, 200), st.integers(1, 4))
def test_rank_to_device_policies_agree_between_cpp_and_python(native, policy, n_devices, rank, n_domains):
    py = tm.device_for_rank(policy, rank, n_devices, None, n_domains)
    cpp = native.topology_device_for_rank(policy, rank, n_devices, [], n_domains)
    assert py == cpp and 0 <= py < n_devices


@FAST
@given(st.sampled_from(["compact", "spread"]), st.sampled_from([1, 2, 4, 6, 8, 12, 16]), st.sampled_from([1, 2]))
def test_first_n_ranks_cover_every_device_once(native, policy, n_devices, n_domains):
    if n_devices % n_domains:
        return
    got = sorted(native.topology_device_for_rank(policy, r, n_devices, [], n_domains) for r in range(n_devices))
    assert got == list(range(n_devices))                      # a permutation: no GPU idle, none shared
    # and the mapping wraps around for oversubscription (cf. devices.hpp round-robin upstream)
    first = native.topology_device_for_rank(policy, 0, n_devices, [], n_domains)
    assert native.topology_device_for_rank(policy, n_devices, n_devices, [], n_domains) == first


@FAST
@given(st.integers(1, 12), st.data())
def test_planes_are_the_connected_components(native, n_gpus, data):
    """Links are sets of endpoint names (GPU ordinals or switch names); planes = connected components of the
    GPUs, members ascending, planes ordered by their first member (topology.cpp:60-90 upstream)."""
    names = [str(g) for g in range(n_gpus)] + ["swA", "swB"]
    links = data.draw(st.lists(st.lists(st.sampled_from(names), min_size=2, max_size=3), max_size=14))
    planes = native.topology_merge_planes(n_gpus, links)
    parent = {n: n for n in names}

    def find(x):
        while parent[x] != x:
            x = parent[x]
        return x

    for link in links:
        for other in link[1:]:
            parent[find(other)] = find(link[0])
    comps = {}
    for g in range(n_gpus):
        comps.setdefault(find(str(g)), []).append(g)
    expected = sorted((sorted(c) for c in comps.values()), key=lambda c: c[0])
    assert planes == expected
    flat = [g for p in planes for g in p]
    for k in range(2 * n_gpus):                                 # compact_plan walks the flattened planes (topology k)
        assert native.topology_device_for_rank("compact_plan", k, n_gpus, planes, 2) == flat[k % n_gpus]
        assert tm.device_for_rank("compact_plan", k, n_gpus, planes) == flat[k % n_gpus]


@FAST
@given(st.floats(1.0, 16.0), st.floats(0.05, 16.0), st.floats(0.0, 4.0))
def test_verdict_rule_and_monotonicity(native, max_speedup, speedup, extra):
    v = native.concurency_judge(max_speedup, speedup, 10.0, -1.0, 100)
    assert v.startswith("FAILURE: Far") == (max_speedup >= 1.3 * speedup)        # TOL_SPEEDUP = 0.3 (main.cpp:314)
    better = native.concurency_judge(max_speedup, speedup + extra, 10.0, -1.0, 100)
    assert not (v.startswith("SUCCESS") and better.startswith("FAILURE"))       # more overlap never hurts
    # the bandwidth floor dominates and only applies when bytes moved
    assert native.concurency_judge(max_speedup, speedup, 1.0, 5.0, 100).startswith("FAILURE: Minimun Bandwish")
    assert native.concurency_judge(max_speedup, speedup, 1.0, 5.0, 0) == native.concurency_judge(max_speedup, speedup, 1.0, -1.0, 0)


@settings(max_examples=60, deadline=None)
@given(st.floats(1e-2, 1.0), st.floats(2e-6, 2e-4), st.floats(2e-6, 2e-4), st.floats(0.0, 1.0),
       st.sampled_from(["in_order", "out_of_order", "nowait", "fused"]))
def test_autotune_balances_linear_commands_and_verdict_follows_overlap(native, c_coef, md_coef, dm_coef, overlap, mode):
    """With a backend whose command times are exactly linear in the tuned parameter, the reference's autotuner
    (main.cpp:219-258: scale every parameter by fastest-copy-time / own-time) makes all commands equally long,
    so the theoretical speedup is the number of commands and the verdict only depends on the overlap.
    (Times are integer microseconds like upstream's; the coefficient ranges keep the baseline runs >= 400 us so
    that the quantisation error of the linear model stays below the 1 % tolerance.)"""
    import re

    spec = f"fake:C={c_coef},MD={md_coef},DM={dm_coef},overlap={overlap}"
    rc, out, err = native.concurency_main([mode, "--commands", "C", "M2D", "D2M"], spec)
    assert "# Performing Autotuning to Balance Commands Times" in out, out + err
    times = [int(t) for t in re.findall(r"Minimum Time Command \d \( *\w+\): (\d+)us", out)]
    assert len(times) == 3 and max(times) - min(times) <= max(2, 0.01 * max(times)), out
    max_speedup = float(re.search(r"Maximum Theoretical Speedup: ([\d.e+-]+)x", out).group(1))
    speedup = float(re.search(r"Speedup Relative to Serial: ([\d.e+-]+)x", out).group(1))
    assert abs(max_speedup - 3.0) < 0.05
    # serial = 3t, concurrent = t + (1 - overlap) * 2t
    assert abs(speedup - 3.0 / (1.0 + 2.0 * (1.0 - overlap))) < 0.05
    ok = "SUCCESS: Close from Theoretical Speedup" in out
    assert ok == (not max_speedup >= 1.3 * speedup) and rc == (0 if ok else 1)
