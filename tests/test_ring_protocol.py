"""Executable model of the fused ring allreduce's flow control (csrc/kernels/ring_allreduce.cu), on the CPU.

The kernel's slot indices and its "who waits for / publishes an ack at which hop" rules live in
csrc/kernels/ring_order.h and are called here through the extension, so the model runs the SAME rules the device
code runs.  P ranks x G CTAs are coroutines that a random scheduler interleaves chunk by chunk; a write into a
slot chunk whose previous content has not been consumed yet is a protocol violation (on the GPU: a data race).
"""
import random

import pytest
from hypothesis import given, settings, strategies as st

import hpc_patterns_b200


class Violation(Exception):
    pass


def simulate(world, n_chunks, ctas, two_slots, seed, honour_acks=True, epoch_base=0):
    C = hpc_patterns_b200.native()
    rng = random.Random(seed)
    n_slots = 2 if two_slots else max(world - 1, 1)
    va = [[(r + 1) * 1000 + c for c in range(n_chunks)] for r in range(world)]
    vc = [[0] * n_chunks for _ in range(world)]
    slots = [[[None] * n_chunks for _ in range(n_slots)] for _ in range(world)]
    unread = [[[False] * n_chunks for _ in range(n_slots)] for _ in range(world)]
    arrived = [[epoch_base] * n_chunks for _ in range(world)]
    ack = [[epoch_base] * n_chunks for _ in range(world)]
    # one coroutine per (rank, cta): position = index into its (t, c) work list (t outer, c inner: kernel order)
    work = {(r, g): [(t, c) for t in range(world) for c in range(g, n_chunks, ctas)]
            for r in range(world) for g in range(ctas)}
    pos = {k: 0 for k in work}

    def runnable(r, t, c):
        if t > 0 and arrived[r][c] < epoch_base + t:
            return False
        if two_slots and honour_acks and C.ring_waits_for_ack(t, world) and ack[r][c] < epoch_base + t - 1:
            return False
        return True

    def step(r, t, c):
        right, left = (r + 1) % world, (r - 1) % world
        if t == 0:
            x = va[r][c]
        else:
            s = C.ring_src_slot(t, two_slots)
            x = slots[r][s][c]
            unread[r][s][c] = False
        vc[r][c] += x
        if C.ring_forwards(t, world):
            s = C.ring_fwd_slot(t, two_slots)
            if unread[right][s][c]:
                raise Violation(f"rank {r} hop {t} overwrites chunk {c} of rank {right}'s slot {s} before it was read")
            slots[right][s][c] = x
            unread[right][s][c] = True
            arrived[right][c] = epoch_base + t + 1
        if two_slots and C.ring_publishes_ack(t, world):
            ack[left][c] = epoch_base + t

    while True:
        ready = [k for k in work if pos[k] < len(work[k]) and runnable(k[0], *work[k][pos[k]])]
        if not ready:
            break
        k = rng.choice(ready)
        step(k[0], *work[k][pos[k]])
        pos[k] += 1
    stuck = [k for k in work if pos[k] < len(work[k])]
    assert not stuck, f"deadlock: {stuck[:4]} blocked"
    total = [sum(va[r][c] for r in range(world)) for c in range(n_chunks)]
    for r in range(world):
        assert vc[r] == total
    return True


@given(world=st.integers(1, 9), n_chunks=st.integers(1, 7), ctas=st.integers(1, 3), two_slots=st.booleans(),
       seed=st.integers(0, 10**6), base=st.sampled_from([0, 40, 2**32 - 3]))
@settings(max_examples=150, deadline=None)
def test_ring_protocol_is_race_free_and_deadlock_free(world, n_chunks, ctas, two_slots, seed, base):
    assert simulate(world, n_chunks, ctas, two_slots, seed, epoch_base=base % (2**31))


def test_model_detects_the_race_when_acks_are_ignored():
    """The checker is not vacuous: with two slots and no flow control some schedule lets a fast sender overwrite
    a chunk its neighbour has not read yet (P >= 4)."""
    hits = 0
    for seed in range(200):
        try:
            simulate(6, 3, 1, True, seed, honour_acks=False)
        except Violation:
            hits += 1
    assert hits > 0


@pytest.mark.parametrize("world", range(1, 10))
def test_ack_rules_pair_up(world):
    """Every ack a sender waits for is published by its neighbour, and nothing else is published."""
    C = hpc_patterns_b200.native()
    waited = {t - 1 for t in range(world) if C.ring_waits_for_ack(t, world)}
    published = {t for t in range(world) if C.ring_publishes_ack(t, world)}
    assert waited == published
    if world <= 3:
        assert not waited


# ------------------------------------------------------------------------------------------------ pull variant ----
def simulate_pull(world, n_chunks, ctas, two_slots, seed, honour_acks=True, epoch_base=0):
    """Receiver-driven ring (ring_pull_kernel): a rank reads hop t-1's block from its LEFT neighbour (its VA for
    t = 1, else the copy the neighbour kept), adds it, and keeps a copy for its own right neighbour.  A copy that is
    overwritten before the right neighbour has read it is a violation."""
    C = hpc_patterns_b200.native()
    rng = random.Random(seed)
    n_slots = 2 if two_slots else max(world - 1, 1)
    va = [[(r + 1) * 1000 + c for c in range(n_chunks)] for r in range(world)]
    vc = [[0] * n_chunks for _ in range(world)]
    copies = [[[None] * n_chunks for _ in range(n_slots)] for _ in range(world)]
    unread = [[[False] * n_chunks for _ in range(n_slots)] for _ in range(world)]
    arrived = [[epoch_base] * n_chunks for _ in range(world)]
    ack = [[epoch_base] * n_chunks for _ in range(world)]
    work = {(r, g): [(t, c) for t in range(world) for c in range(g, n_chunks, ctas)]
            for r in range(world) for g in range(ctas)}
    pos = {k: 0 for k in work}

    def runnable(r, t, c):
        if t > 0 and arrived[r][c] < epoch_base + t:
            return False
        if two_slots and honour_acks and C.ring_pull_waits_for_ack(t, world) and ack[r][c] < epoch_base + t - 1:
            return False
        return True

    def step(r, t, c):
        right, left = (r + 1) % world, (r - 1) % world
        if t == 0:
            x = va[r][c]
        elif t == 1:
            x = va[left][c]
        else:
            s = C.ring_pull_src_slot(t, two_slots)
            x = copies[left][s][c]
            if x is None or not unread[left][s][c]:
                raise Violation(f"rank {r} hop {t} reads chunk {c} of rank {left}'s slot {s} which does not hold hop {t - 1}")
            unread[left][s][c] = False
        vc[r][c] += x
        if C.ring_pull_keeps_copy(t, world):
            s = C.ring_pull_copy_slot(t, two_slots)
            if unread[r][s][c]:
                raise Violation(f"rank {r} hop {t} overwrites its copy slot {s} chunk {c} before rank {right} read it")
            copies[r][s][c] = x
            unread[r][s][c] = True
        if t + 1 < world:
            arrived[right][c] = epoch_base + t + 1
        if two_slots and C.ring_pull_publishes_ack(t, world):
            ack[left][c] = epoch_base + t

    while True:
        ready = [k for k in work if pos[k] < len(work[k]) and runnable(k[0], *work[k][pos[k]])]
        if not ready:
            break
        k = rng.choice(ready)
        step(k[0], *work[k][pos[k]])
        pos[k] += 1
    stuck = [k for k in work if pos[k] < len(work[k])]
    assert not stuck, f"deadlock: {stuck[:4]} blocked"
    total = [sum(va[r][c] for r in range(world)) for c in range(n_chunks)]
    for r in range(world):
        assert vc[r] == total
    return True


@given(world=st.integers(1, 9), n_chunks=st.integers(1, 7), ctas=st.integers(1, 3), two_slots=st.booleans(),
       seed=st.integers(0, 10**6), base=st.sampled_from([0, 40]))
@settings(max_examples=150, deadline=None)
def test_pull_ring_protocol_is_race_free_and_deadlock_free(world, n_chunks, ctas, two_slots, seed, base):
    assert simulate_pull(world, n_chunks, ctas, two_slots, seed, epoch_base=base)


def test_pull_model_detects_the_race_when_acks_are_ignored():
    hits = 0
    for seed in range(300):
        try:
            simulate_pull(7, 3, 1, True, seed, honour_acks=False)
        except Violation:
            hits += 1
    assert hits > 0


@pytest.mark.parametrize("world", range(1, 10))
def test_pull_ack_rules_pair_up(world):
    C = hpc_patterns_b200.native()
    waited = {t - 1 for t in range(world) if C.ring_pull_waits_for_ack(t, world)}
    published = {t for t in range(world) if C.ring_pull_publishes_ack(t, world)}
    assert waited == published
    if world <= 4:
        assert not waited                       # with P <= 4 no copy is ever overwritten inside one allreduce
    for t in range(2, world):                   # what a rank keeps at hop t-1 is what its neighbour reads at hop t
        for two in (False, True):
            assert C.ring_pull_src_slot(t, two) == C.ring_pull_copy_slot(t - 1, two)
