"""The flagship's Python layer (models/halo.py, parallel/local.py, parallel/symmetric.py) on the CPU, against an
EMULATED native module.

Everything those classes do goes through ``native()``: allocate, copy, launch.  ``EmuNative`` implements that surface on
host memory — addresses are real host addresses, a "launch" executes at once — with the semantics of the kernels it
stands in for: `halo_stencil` reads exactly the addresses it is handed (the neighbours' fields in pull mode, the local
halo buffers otherwise), steps the column range it is told to, stores where it is told to (push: also into the
neighbours' halo buffers), REFUSES to run if the step words it would spin on are not there yet (a launch that would wait
for something nobody enqueued is a deadlock of the one-thread driver), and publishes its own.  So the real
`HaloStencil` / `VirtualRing` / `SymmetricBuffer` / `SignalPads` code runs unmodified: every parity, row offset, halo
side, flag pointer, flag set and chunk boundary it computes is checked by comparing the resulting field with a plain
PyTorch run of the undecomposed stencil, bit for bit.  The GPU suite does the same with the real kernels
(tests/test_gpu_halo.py); this runs on the CPU-only box of every round.
"""
import pytest
import torch

from hpc_patterns_b200.models import halo as halo_mod
from hpc_patterns_b200.models.halo import initial_field, reference_steps
from tests.emu_device import EmuNative, install


@pytest.fixture
def emu(monkeypatch):
    """The real classes on an emulated device: native() -> EmuNative, torch.cuda.* -> no-ops, 'cuda' tensors -> CPU."""
    e = EmuNative()
    install(monkeypatch.setattr, e)
    return e


def _want(world, rows, n, steps):
    return reference_steps(initial_field(world, rows, n), steps)


@pytest.mark.parametrize("mode", ["pull", "push"])
@pytest.mark.parametrize("world,rows,n", [(1, 1, 64), (2, 3, 5000), (3, 2, 8192 + 4), (5, 4, 4100)])
def test_virtual_ring_pointer_and_parity_plumbing(emu, mode, world, rows, n):
    ring = halo_mod.VirtualRing(world, 4 * n, rows, mode, devices=[0] * world, tune={"tile_kb": 4})
    ring.step(5)
    assert torch.equal(ring.gather(), _want(world, rows, n, 5))
    for hs in ring.ranks:
        assert hs.g == 5 and hs.launches == 5
        assert hs.verify_from_init() == 0 and hs.verify_last_step() == 0
    ring.close()
    assert not emu.live, "every allocation is returned"


def test_a_rank_ahead_of_its_neighbours_is_refused(emu):
    """The step words are real: stepping ONE rank of a ring twice asks for a step its neighbours never produced."""
    ring = halo_mod.VirtualRing(3, 4 * 1024, 2, "pull", devices=[0, 0, 0], tune={"tile_kb": 4})
    ring.ranks[0].step(1)
    with pytest.raises(RuntimeError, match="nobody has enqueued"):
        ring.ranks[0].step(1)


def test_single_rank_persistent_launch_and_reset(emu):
    hs = halo_mod.VirtualRing(1, 4 * 3000, 3, "push", devices=[0], tune={"tile_kb": 4}).ranks[0]
    hs.step(7)                                     # one launch, seven steps: the rank is its own neighbour
    assert hs.launches == 1 and torch.equal(hs.u_tensor(), _want(1, 3, 3000, 7))
    hs.reset()
    assert hs.g == 0 and torch.equal(hs.u_tensor(), initial_field(1, 3, 3000))
    hs.step(2)
    assert torch.equal(hs.u_tensor(), _want(1, 3, 3000, 2))


def test_stock_steps_are_the_same_time_series_and_do_not_mix_with_fused_steps(emu):
    ring = halo_mod.VirtualRing(1, 4 * 2048, 2, "pull", devices=[0], tune={"tile_kb": 4})
    hs = ring.ranks[0]
    for _ in range(4):
        hs.stock_step("memcpy")
    assert torch.equal(hs.u_tensor(), _want(1, 2, 2048, 4)) and hs.verify_last_step() == 0
    with pytest.raises(RuntimeError, match="reset"):
        hs.step(1)                                 # the stock steps never advanced the step words
    hs.reset()
    hs.step(3)
    assert torch.equal(hs.u_tensor(), _want(1, 2, 2048, 3))
    assert [l[0] for l in emu.launches[:4]] == ["none"] * 4


def test_exchange_only_moves_the_two_boundary_rows(emu):
    hs = halo_mod.VirtualRing(1, 4 * 1024, 3, "push", devices=[0], tune={"tile_kb": 4}).ranks[0]
    hs.step(2)
    par = (hs.g + 1) & 1
    rows = hs.u_tensor(par).clone()                # what exchange_only sends: rows 0 and R-1 of the NEXT parity
    hs.exchange_only()
    hi = halo_mod.tensor_from_ptr(hs.halo_ptr(0, "hi", par), hs.row_bytes, 0, torch.float32)
    lo = halo_mod.tensor_from_ptr(hs.halo_ptr(0, "lo", par), hs.row_bytes, 0, torch.float32)
    assert torch.equal(hi, rows[0]) and torch.equal(lo, rows[2])


@pytest.mark.parametrize("world,chunks,n", [(1, 16, 3 * 1024 + 8), (2, 4, 8 * 1024), (3, 3, 5 * 1024 + 256),
                                            (2, 1, 2048)])
def test_out_of_core_steps_through_host_memory(emu, world, chunks, n):
    """step_from_host: the slab lives in (pinned) host memory, every step uploads it in column chunks, steps each
    chunk with its own flag set and downloads the result; halos travel rank to rank (push)."""
    rows = 3
    ring = halo_mod.VirtualRing(world, 4 * n, rows, "push", devices=[0] * world, tune={"tile_kb": 1})
    bufs = [hs.make_host_buffers() for hs in ring.ranks]
    for k in range(4):
        for hs, b in zip(ring.ranks, bufs):
            b[(k + 1) & 1].fill_(float("nan"))
            hs.step_from_host(b[k & 1], b[(k + 1) & 1], chunks=chunks)
    got = torch.cat([b[0] for b in bufs], 0)       # four steps: the result is back in buffer 0
    assert torch.equal(got, _want(world, rows, n, 4))
    used_sets = {l[5] for l in emu.launches}
    assert used_sets == set(range(min(chunks, n * 4 // 1024 + (1 if (n * 4) % 1024 else 0), emu.HALO_FLAG_SETS)))
    if len(used_sets) > 1:
        with pytest.raises(RuntimeError, match="reset"):
            ring.ranks[0].step(1)                  # chunked and whole-row steps keep different step words
    else:                                          # one chunk = the whole-row layout: the series simply continues
        for hs in ring.ranks:
            hs.step(1)
        assert torch.equal(ring.gather(), _want(world, rows, n, 5))


def test_step_from_host_needs_push(emu):
    hs = halo_mod.VirtualRing(1, 4 * 1024, 2, "pull", devices=[0], tune={"tile_kb": 4}).ranks[0]
    b = hs.make_host_buffers()
    with pytest.raises(RuntimeError, match="push"):
        hs.step_from_host(b[0], b[1])


@pytest.mark.parametrize("mode,mutation", [("push", "put_targets"), ("pull", "parity"), ("push", "flags")])
def test_the_emulation_notices_wrong_plumbing(emu, monkeypatch, mode, mutation):
    """Sensitivity check of this file: put targets the wrong way round, a stuck parity or a step word published to the
    wrong rank in the Python layer must not survive the comparison."""
    if mutation == "parity":
        real_u = halo_mod.HaloStencil.u_ptr
        monkeypatch.setattr(halo_mod.HaloStencil, "u_ptr",
                            lambda self, rank, parity, row=0: real_u(self, rank, parity if rank == self.rank else 0, row))
    else:
        real_args = halo_mod.HaloStencil._args

        def mutated(self, *a, **k):
            d = real_args(self, *a, **k)
            if mutation == "put_targets":
                d["left_halo_hi"], d["right_halo_lo"] = d["right_halo_lo"], d["left_halo_hi"]
            elif self.rank == 0:
                d["flags_left"] = d["flags_right"]      # rank 0's "I am done" never reaches its left neighbour
            return d
        monkeypatch.setattr(halo_mod.HaloStencil, "_args", mutated)
    ring = halo_mod.VirtualRing(3, 4 * 2048, 2, mode, devices=[0, 0, 0], tune={"tile_kb": 4})
    try:
        ring.step(3)
        wrong = not torch.equal(ring.gather(), _want(3, 2, 2048, 3))
    except RuntimeError as e:           # the flag mutation shows up as a launch that would wait for ever
        wrong = "nobody has enqueued" in str(e)
    assert wrong
