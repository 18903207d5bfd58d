import json
import os
import subprocess

import pytest

from hpc_patterns_b200.parallel import tile_mapping as tm


def test_merge_planes_pvc_like_two_planes(native):
    # six "tiles": two fully connected planes, links listed in a scrambled order
    links = [["0", "2"], ["3", "5"], ["2", "4"], ["1", "3"], ["0", "4"], ["5", "1"]]
    assert native.topology_merge_planes(6, links) == [[0, 2, 4], [1, 3, 5]]


def test_merge_planes_transitive_and_isolated(native):
    assert native.topology_merge_planes(5, [["0", "1"], ["1", "2"]]) == [[0, 1, 2], [3], [4]]


def test_nvswitch_is_one_plane(native):
    links = [[str(g), "nvswitch"] for g in range(8)]
    assert native.topology_merge_planes(8, links) == [list(range(8))]
    info = json.loads(native.topology_discover("8:switch"))
    assert info["planes"] == [list(range(8))] and info["source"] == "fake"
    assert all(g["nvlinks_to_switch"] == 18 for g in info["gpus"])


def test_two_switch_domains(native):
    info = json.loads(native.topology_discover("8:switch:0-3;4-7"))
    assert info["planes"] == [[0, 1, 2, 3], [4, 5, 6, 7]]


@pytest.mark.parametrize("policy,expected", [
    ("compact", [0, 1, 2, 3, 4, 5, 6, 7]),
    ("spread", [0, 4, 1, 5, 2, 6, 3, 7]),
])
def test_policies_8_gpus(native, policy, expected):
    got = [native.topology_device_for_rank(policy, r, 8) for r in range(8)]
    assert got == expected
    assert [tm.device_for_rank(policy, r, 8) for r in range(8)] == expected   # python mirror agrees


def test_compact_plan_pairs_share_a_plane(native):
    planes = [[0, 2, 4], [1, 3, 5]]
    got = [native.topology_device_for_rank("compact_plan", r, 6, planes) for r in range(6)]
    assert got == [0, 2, 4, 1, 3, 5]
    assert [tm.device_for_rank("compact_plan", r, 6, planes) for r in range(6)] == got
    flat_planes = {g: i for i, p in enumerate(planes) for g in p}
    assert flat_planes[got[0]] == flat_planes[got[1]]


def test_oversubscription_wraps(native):
    assert [native.topology_device_for_rank("compact", r, 2) for r in range(5)] == [0, 1, 0, 1, 0]
    with pytest.raises(Exception):
        native.topology_device_for_rank("nonsense", 0, 2)


def test_topology_cli(bin_dir):
    exe = os.path.join(bin_dir, "topology")
    env = dict(os.environ, HPCP_FAKE_TOPOLOGY="6:0-2,2-4,0-4,1-3,3-5,1-5")
    out = subprocess.run([exe], env=env, capture_output=True, text=True).stdout
    assert out.splitlines() == ["0 2 4 ", "1 3 5 "]
    assert subprocess.run([exe, "3"], env=env, capture_output=True, text=True).stdout.strip() == "1"
    assert subprocess.run([exe, "--policy", "spread", "--rank", "1", "--ndev", "6"], env=env,
                          capture_output=True, text=True).stdout.strip() == "3"
    p = subprocess.run([exe, "99"], env=env, capture_output=True, text=True)
    assert p.returncode == 1
    j = json.loads(subprocess.run([exe, "--json"], env=env, capture_output=True, text=True).stdout)
    assert j["planes"] == [[0, 2, 4], [1, 3, 5]]


def test_topology_cli_without_driver_fails_cleanly(bin_dir, gpu_count):
    if gpu_count:
        pytest.skip("GPU present")
    env = {k: v for k, v in os.environ.items() if k != "HPCP_FAKE_TOPOLOGY"}
    p = subprocess.run([os.path.join(bin_dir, "topology")], env=env, capture_output=True, text=True)
    assert p.returncode == 1 and "no fabric information" in p.stderr
