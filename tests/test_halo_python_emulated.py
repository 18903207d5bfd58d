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
import contextlib
import ctypes

import numpy as np
import pytest
import torch

import hpc_patterns_b200
from hpc_patterns_b200.models import halo as halo_mod
from hpc_patterns_b200.models.halo import initial_field, reference_steps
from hpc_patterns_b200.parallel import local as local_mod
from hpc_patterns_b200.parallel import symmetric as symmetric_mod

FLAG_WORDS = 8          # kHaloFlagWords (csrc/kernels/api.h): one 32-byte sector per CTA


def _f32(ptr, n):
    return torch.frombuffer((ctypes.c_float * n).from_address(ptr), dtype=torch.float32)


def _u32(ptr):
    return ctypes.c_uint32.from_address(ptr)


class EmuNative:
    def __init__(self, ctas=3):
        self.real = hpc_patterns_b200.native()
        self.ctas = ctas
        self.live = {}
        self.launches = []

    def __getattr__(self, name):          # constants (PAD_*, HALO_*, STATUS_*) come from the real extension
        attr = getattr(self.real, name)
        if callable(attr):
            raise AttributeError(f"EmuNative: {name} is not emulated")
        return attr

    # ---- memory ----------------------------------------------------------------------------
    def alloc(self, nbytes, kind="D", device=0, zero=True):
        buf = np.zeros(nbytes + 64, dtype=np.uint8) if zero else np.full(nbytes + 64, 0xA5, dtype=np.uint8)
        addr = (buf.ctypes.data + 63) & ~63
        self.live[addr] = buf
        return addr

    def free(self, ptr, kind="D"):
        del self.live[ptr]

    def memset_async(self, ptr, value, nbytes, stream):
        ctypes.memset(ptr, value, nbytes)

    def memcpy_async(self, dst, src, nbytes, stream):
        ctypes.memmove(dst, src, nbytes)

    def read_u32(self, ptr):
        return _u32(ptr).value

    def enable_peer_access(self, devices):
        pass

    # ---- flags -------------------------------------------------------------------------------
    def signal(self, flag, epoch, stream):
        _u32(flag).value = epoch

    def wait(self, flag, epoch, timeout_ns, status, stream):
        if _u32(flag).value < epoch:
            raise RuntimeError(f"wait for epoch {epoch} would never return: the word holds {_u32(flag).value}")

    def copy(self, dst, src, nbytes, src_is_peer=False, engine="ldst", tune=None, sync=None, device=0, stream=0):
        sync = sync or {}
        if "wait_flag" in sync:
            self.wait(sync["wait_flag"], sync["wait_epoch"], 0, 0, 0)
        ctypes.memmove(dst, src, nbytes)
        if "signal_flag" in sync:
            _u32(sync["signal_flag"]).value = sync["signal_epoch"]
        return 2

    # ---- K-halo ------------------------------------------------------------------------------
    def halo_stencil_ctas(self, row_elems, mode="pull", tune=None, device=0):
        return int((tune or {}).get("ctas") or self.ctas)

    def halo_init(self, u, halo_lo, halo_hi, rows, row_elems, rank, world, stream):
        f = initial_field(world, rows, row_elems)
        G, first = world * rows, rank * rows
        _f32(u, rows * row_elems).view(rows, row_elems).copy_(f[first:first + rows])
        if halo_lo:
            _f32(halo_lo, row_elems).copy_(f[(first + G - 1) % G])
        if halo_hi:
            _f32(halo_hi, row_elems).copy_(f[(first + rows) % G])

    def _flag(self, base, flag_set, side, cta):
        return base + 4 * ((flag_set * 2 + side) * self.real.HALO_MAX_CTAS + cta) * FLAG_WORDS

    def halo_stencil(self, a, mode="pull", tune=None, device=0, stream=0):
        tune = tune or {}
        R, n = a["rows"], a["row_elems"]
        tile = (tune.get("tile_kb") or 16) * 1024 // 4
        tiles = (n + tile - 1) // tile
        t0, t1 = a.get("tile_begin", 0), a.get("tile_end", 0) or tiles
        assert 0 <= t0 < t1 <= tiles, "bad column-tile range"
        c0, c1 = t0 * tile, min(n, t1 * tile)
        ctas = self.halo_stencil_ctas(n, mode, tune)
        alpha = torch.tensor(a["alpha"], dtype=torch.float32)
        s = torch.tensor(a["s"], dtype=torch.float32)
        fs = a.get("flag_set", 0)
        self.launches.append((mode, a["step_base"], a["steps"], t0, t1, fs))
        if mode != "none" and a["steps"] > 1:
            own = a["left_u"][0] == a["u"][0] and a["right_u"][0] == a["u"][0]
            assert own, "the emulator runs a multi-step launch only when the rank is its own neighbour"
        for g in range(a["step_base"], a["step_base"] + a["steps"]):
            i, o = g & 1, (g + 1) & 1
            if mode != "none":       # the DMA thread's wait_epoch on both neighbour words of every CTA
                for c in range(ctas):
                    for side in (0, 1):
                        have = _u32(self._flag(a["flags_local"], fs, side, c)).value
                        if have < g:
                            raise RuntimeError(f"step {g}: word (set {fs}, side {side}, cta {c}) holds {have}: "
                                               "this launch would spin on a step nobody has enqueued")
            u_in = _f32(a["u"][i], R * n).view(R, n)
            u_out = _f32(a["u"][o], R * n).view(R, n)
            if mode == "pull":
                up = _f32(a["left_u"][i], R * n).view(R, n)[R - 1]
                dn = _f32(a["right_u"][i], R * n).view(R, n)[0]
            else:
                up, dn = _f32(a["halo_lo"][i], n), _f32(a["halo_hi"][i], n)
            ext = torch.cat([up[None, c0:c1], u_in[:, c0:c1], dn[None, c0:c1]], 0)
            new = alpha * ext[1:-1] + s * (ext[:-2] + ext[2:])
            u_out[:, c0:c1] = new
            if mode == "push":
                _f32(a["left_halo_hi"][o], n)[c0:c1] = new[0]
                _f32(a["right_halo_lo"][o], n)[c0:c1] = new[R - 1]
            if mode != "none":       # st.release.sys of g + 1 on both neighbours, per CTA
                for c in range(ctas):
                    _u32(self._flag(a["flags_right"], fs, 0, c)).value = g + 1   # its wait_lo: I am its left neighbour
                    _u32(self._flag(a["flags_left"], fs, 1, c)).value = g + 1    # its wait_hi
        return ctas

    def halo_verify_from_init(self, u, rows, row_elems, rank, world, steps, alpha, s, mismatch, stream):
        want = reference_steps(initial_field(world, rows, row_elems), steps, alpha, s)[rank * rows:(rank + 1) * rows]
        got = _f32(u, rows * row_elems).view(rows, row_elems)
        ctypes.c_int64.from_address(mismatch).value += int((got != want).sum())

    def halo_verify_step(self, u_new, u_old, up_row, dn_row, rows, row_elems, alpha, s, mismatch, stream):
        old = _f32(u_old, rows * row_elems).view(rows, row_elems)
        ext = torch.cat([_f32(up_row, row_elems)[None], old, _f32(dn_row, row_elems)[None]], 0)
        a, b = torch.tensor(alpha, dtype=torch.float32), torch.tensor(s, dtype=torch.float32)
        want = a * ext[1:-1] + b * (ext[:-2] + ext[2:])
        got = _f32(u_new, rows * row_elems).view(rows, row_elems)
        ctypes.c_int64.from_address(mismatch).value += int((got != want).sum())


class _Stream:
    cuda_stream = 0

    def __init__(self, *a, **k):
        pass

    def synchronize(self):
        pass

    def wait_stream(self, other):
        pass

    def wait_event(self, event):
        pass


class _Event:
    def __init__(self, *a, **k):
        pass

    def record(self, stream=None):
        pass


@pytest.fixture
def emu(monkeypatch):
    """The real classes on an emulated device: native() -> EmuNative, torch.cuda.* -> no-ops, 'cuda' tensors -> CPU."""
    e = EmuNative()
    for mod in (halo_mod, local_mod, symmetric_mod):
        monkeypatch.setattr(mod, "native", lambda: e)
    monkeypatch.setattr(halo_mod, "tensor_from_ptr",
                        lambda ptr, nbytes, device, dtype=torch.uint8:
                        torch.frombuffer((ctypes.c_uint8 * nbytes).from_address(ptr), dtype=torch.uint8).view(dtype))
    for name, fake in (("set_device", lambda d=None: None), ("synchronize", lambda d=None: None),
                       ("current_stream", lambda d=None: _Stream()), ("Stream", _Stream), ("Event", _Event),
                       ("device", lambda d=None: contextlib.nullcontext()),
                       ("stream", lambda s=None: contextlib.nullcontext()), ("device_count", lambda: 1)):
        monkeypatch.setattr(torch.cuda, name, fake)
    real_zeros = torch.zeros
    monkeypatch.setattr(torch, "zeros", lambda *a, **k: real_zeros(*a, **{**k, "device": "cpu"}))
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
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
