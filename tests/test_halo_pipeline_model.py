"""Executable model of the shared-memory pipeline inside one K-halo CTA (csrc/kernels/halo_stencil.cu), on the CPU.

The kernel keeps ONE ring of S tiles per CTA: the DMA thread streams the R + 2 row tiles of a column tile into
consecutive slots, the math warps write the result of row r over the tile of row r - 1 IN PLACE, and the DMA thread
sends that slot out through the TMA unit; a slot is reloaded as soon as `retired` says its last reader is done.  The
slot cursors advance by 1 inside a column and by 3 at its end (the last centre tile and the lower halo are never
stored), loads may run up to S tiles ahead of the stores, stores complete out of the DMA thread's sight except for
`bulk_wait_read<1>`.  This model restates exactly those rules — the `Ring` cursors, `issued < retired + S`,
`retired = i0` after `wait_read<1>`, the `full` / `computed` barrier counts and parities — and runs the DMA thread
and the math warps as coroutines under a random scheduler that delays every asynchronous completion (bulk loads
landing, bulk stores reading their slot) arbitrarily.  Checked under every schedule: no deadlock, no arrival on a
completed barrier phase, every stencil reads the three tiles it is meant to, every store sends the result it is meant
to, and no load lands in a slot that a store is still reading or that the math warps still need.  Mutated rules (a slot
freed one store too early, a ring shorter than the pipeline needs, a wrong end-of-column skip) must be caught.
"""
import random

import pytest

from tests.test_umma_pipeline_model import MBar, ProtocolError, run

MATH_WARPS = 4


class Cursor:
    """`struct Ring` of the kernel: slot index + phase parity of its current use."""

    def __init__(self, stages):
        self.slot, self.phase, self.S = 0, 0, stages

    def advance(self, n):
        self.slot += n
        if self.slot >= self.S:
            self.slot -= self.S
            self.phase ^= 1

    def at(self, n):
        c = Cursor(self.S)
        c.slot, c.phase = self.slot, self.phase
        c.advance(n)
        return c


def simulate(stages, rows, ncols, steps, seed, wait_read_keep=1, free_extra=0, column_skip=3):
    """wait_read_keep: N of `bulk_wait_read<N>`; free_extra: tiles freed beyond `retired = i0`; column_skip: cursor
    advance at the end of a column.  The defaults are the kernel's."""
    rng = random.Random(seed)
    S, R = stages, rows
    full = [MBar(f"full[{i}]", 1) for i in range(S)]
    computed = [MBar(f"computed[{i}]", MATH_WARPS) for i in range(S)]
    # slot contents: ("in", step, col, m) for an input tile (m = row + 1), ("out", step, col, r) for a result
    content = [None] * S
    busy_store = [0] * S          # bulk stores that have been issued on the slot and have not read it yet
    needed_by_math = [0] * S      # (computation, warp) pairs that still have to read the slot's current tile
    writes = [0] * S              # warps that have written their quarter of the result into the slot
    load_q, store_q = [], []      # asynchronous completions, FIFO per engine queue
    stored = []                   # what left the CTA, in completion order
    loads_per_step, comps_per_step = ncols * (R + 2), ncols * R

    def dma():
        ld, st, done = Cursor(S), Cursor(S), Cursor(S)
        for g in range(steps):
            issued = retired = 0
            jj_l = m_l = 0
            jj_c = r = i0 = 0
            outstanding = []      # stores issued and not yet known to have read their slot, oldest first
            for _ in range(comps_per_step):
                while issued < loads_per_step and issued < retired + S + free_extra:
                    slot = ld.slot
                    full[slot].arrive(expect_tx=1)
                    # how many stencils will read this tile: row m-1 as "down", row m as "centre", row m+1 as "up"
                    readers = sum(1 for rr in (m_l - 2, m_l - 1, m_l) if 0 <= rr < R)

                    def landed(slot=slot, tile=("in", g, jj_l, m_l), readers=readers):
                        if busy_store[slot]:
                            raise ProtocolError(f"load of {tile} landed in slot {slot} while a store still reads it")
                        if needed_by_math[slot]:
                            raise ProtocolError(f"load of {tile} landed in slot {slot}: {content[slot]} is still needed")
                        content[slot] = tile
                        needed_by_math[slot] = readers * MATH_WARPS
                        full[slot].complete_tx(1)

                    load_q.append(landed)
                    ld.advance(1)
                    issued += 1
                    m_l += 1
                    if m_l == R + 2:
                        m_l, jj_l = 0, jj_l + 1
                yield lambda s=done.slot, p=done.phase: computed[s].done(p)
                done.advance(1)
                slot = st.slot
                busy_store[slot] += 1
                token = [False]

                def read_out(slot=slot, want=("out", g, jj_c, r), token=token):
                    if content[slot] != want:
                        raise ProtocolError(f"store of {want} read {content[slot]} from slot {slot}")
                    stored.append(want)
                    busy_store[slot] -= 1
                    token[0] = True

                store_q.append(read_out)
                outstanding.append(token)
                # cp.async.bulk.wait_group.read N: all but the N newest groups have read their source
                while len(outstanding) > wait_read_keep:
                    oldest = outstanding.pop(0)
                    yield lambda t=oldest: t[0]
                retired = i0
                if r + 1 == R:
                    r, jj_c, i0 = 0, jj_c + 1, i0 + column_skip
                    st.advance(column_skip)
                else:
                    r, i0 = r + 1, i0 + 1
                    st.advance(1)
            for t in outstanding:     # bulk_wait<0> at the end of the step
                yield lambda t=t: t[0]

    def math(warp):
        up_t, done = Cursor(S), Cursor(S)
        for g in range(steps):
            jj = r = 0
            for _ in range(comps_per_step):
                ce_t, dn_t = up_t.at(1), up_t.at(2)
                if r == 0:
                    yield lambda c=up_t: full[c.slot].done(c.phase)
                    yield lambda c=ce_t: full[c.slot].done(c.phase)
                yield lambda c=dn_t: full[c.slot].done(c.phase)
                # row r reads the tiles of rows r-1, r, r+1 (load slots m = r, r+1, r+2 of the column) and writes its result
                # over the first of them; every warp handles its own quarter of the tile
                want = tuple(("in", g, jj, r + k) for k in range(3))
                got = (content[up_t.slot], content[ce_t.slot], content[dn_t.slot])
                if got != want:
                    raise ProtocolError(f"stencil of step {g} col {jj} row {r} (warp {warp}) read {got}, wanted {want}")
                for c in (up_t, ce_t, dn_t):
                    needed_by_math[c.slot] -= 1
                writes[up_t.slot] += 1
                if writes[up_t.slot] == MATH_WARPS:
                    writes[up_t.slot] = 0
                    content[up_t.slot] = ("out", g, jj, r)
                yield None
                computed[done.slot].arrive()
                done.advance(1)
                if r + 1 == R:
                    r, jj = 0, jj + 1
                    up_t.advance(column_skip)
                else:
                    r += 1
                    up_t.advance(1)

    agents = [dma()] + [math(w) for w in range(MATH_WARPS)]
    run(agents, [load_q, store_q], rng)
    want = [("out", g, c, r) for g in range(steps) for c in range(ncols) for r in range(R)]
    if sorted(stored) != sorted(want):
        raise ProtocolError("not every result tile was stored exactly once")
    return len(stored)


@pytest.mark.parametrize("stages", [6, 7, 12])
@pytest.mark.parametrize("rows,ncols", [(1, 3), (2, 2), (3, 4), (8, 2)])
def test_kernel_rules_are_safe_under_random_schedules(stages, rows, ncols):
    for seed in range(25):
        assert simulate(stages, rows, ncols, steps=3, seed=seed) == 3 * rows * ncols


def test_freeing_a_slot_before_its_store_was_read_is_caught():
    """`bulk_wait_read<2>` with `retired = i0` frees the slot of the second-newest store while the TMA unit may still
    be reading it."""
    with pytest.raises(ProtocolError):
        for seed in range(200):
            simulate(6, 3, 4, steps=2, seed=seed, wait_read_keep=2, free_extra=1)


def test_loads_running_too_far_ahead_are_caught():
    """`issued < retired + S + 1`: one tile more than the ring holds."""
    with pytest.raises(ProtocolError):
        for seed in range(200):
            simulate(6, 3, 4, steps=2, seed=seed, free_extra=1)


def test_wrong_end_of_column_skip_is_caught():
    """Advancing by 2 at the end of a column points the next column's cursors at the previous lower-halo tile."""
    with pytest.raises(ProtocolError):
        for seed in range(50):
            simulate(6, 3, 4, steps=2, seed=seed, column_skip=2)


@pytest.mark.parametrize("stages", [3, 4, 5])
def test_ring_shorter_than_six_slots_deadlocks(stages):
    """The first stencil of a column needs its three tiles while `retired` still points at the last STORED tile of the
    previous column (its two never-stored tiles only count as free one computation later): 3 + 2 + the slot being
    stored = 6.  halo_geometry() refuses fewer stages for exactly this reason."""
    with pytest.raises(ProtocolError, match="deadlock"):
        simulate(stages, 3, 4, steps=2, seed=0)
