"""Dry run (no GPU) of how the `-a` path of the allreduce miniapp picks its collective
(models/allreduce.py::choose_collective) and of the failure path of the NVLS constructor: the decision is taken BEFORE
anything is allocated, it is the minimum over the ranks' probes (ranks can not diverge on the algorithm), integers go to
the vector-load two-shot, and a constructor that fails after the collective pad allocation closes the pads."""
import types

import pytest

from hpc_patterns_b200.models import allreduce as ar


class FakeComm:
    def __init__(self, rank, world, others_ok=True):
        self.rank, self.world, self.others_ok = rank, world, others_ok
        self.min_calls = []

    def min(self, v):
        self.min_calls.append(v)
        return min(v, 1.0 if self.others_ok else 0.0)


def _native(multicast):
    return lambda: types.SimpleNamespace(multicast_supported=lambda device: multicast)


def test_float_takes_the_switch_when_every_rank_can(monkeypatch):
    monkeypatch.setattr(ar, "native", _native(True))
    c = FakeComm(0, 8)
    algo, why = ar.choose_collective(c, 0, "float", 25)
    assert algo == "nvls" and c.min_calls == [1.0] and "every rank" in why


def test_one_rank_without_multicast_moves_everybody_to_two_shot(monkeypatch):
    monkeypatch.setattr(ar, "native", _native(True))
    algo, why = ar.choose_collective(FakeComm(0, 8, others_ok=False), 0, "float", 25)
    assert algo == "twoshot" and "at least one rank" in why
    monkeypatch.setattr(ar, "native", _native(False))          # ... and the rank that lacks it agrees
    c = FakeComm(3, 8)
    assert ar.choose_collective(c, 0, "float", 25)[0] == "twoshot" and c.min_calls == [0.0]


@pytest.mark.parametrize("dtype", ["int", "uint", "double", "long", "short", "uchar"])
def test_other_types_use_vector_peer_loads(monkeypatch, dtype):
    monkeypatch.setattr(ar, "native", _native(True))
    c = FakeComm(0, 8)
    algo, _ = ar.choose_collective(c, 0, dtype, 25)
    assert algo == "twoshot" and c.min_calls == []             # decided locally, identically on every rank


def test_single_gpu_has_no_switch_to_reduce_in(monkeypatch):
    monkeypatch.setattr(ar, "native", _native(True))
    assert ar.choose_collective(FakeComm(0, 1), 0, "float", 25)[0] == "twoshot"


def test_failed_nvls_constructor_returns_its_pads(monkeypatch):
    """AllreduceMiniapp(..., 'nvls') allocates the signal pads collectively, then maps the multicast object; if that
    raises, the pads must be closed before the exception leaves (no leak, nothing half-built to fall back from)."""
    closed = []

    class Pads:
        def __init__(self, comm, device, extra_words=0, timeout_s=20.0):
            pass

        def close(self):
            closed.append(True)

    fake_native = types.SimpleNamespace(elem_size=lambda t: 4, ring_num_chunks=lambda n, chunk, esz=4: 1)
    monkeypatch.setattr(ar, "native", lambda: fake_native)
    monkeypatch.setattr(ar, "SignalPads", Pads)
    monkeypatch.setattr(ar.torch.cuda, "set_device", lambda d: None)

    def boom(self):
        raise RuntimeError("torch symmetric memory reports no multicast support on this system")
    monkeypatch.setattr(ar.AllreduceMiniapp, "_init_nvls", boom)
    with pytest.raises(RuntimeError, match="multicast"):
        ar.AllreduceMiniapp(FakeComm(0, 2), 0, 10, "float", "nvls")
    assert closed == [True]
