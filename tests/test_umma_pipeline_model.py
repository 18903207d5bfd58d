"""Executable model of the mbarrier pipelines of the tensor-core tile loops (csrc/kernels/umma.cuh), on the CPU.

The 2-SM UMMA loop (``gemm_persistent_2sm``) was written without access to a GPU; a protocol error there is a hang.
This model re-states its barrier protocol — who initialises which barrier with which count, who arrives / expects
bytes / waits on which parity — with mbarrier semantics (pending arrivals, transaction bytes, phase bit), runs the
warps of a CTA pair as coroutines under a random scheduler in which every asynchronous completion (TMA bytes landing,
tcgen05.commit arrivals) is delayed arbitrarily, and checks: no deadlock, no arrival on a barrier whose phase is
already full, no smem stage refilled before the MMAs that read it were committed, no accumulator overwritten before
both epilogues drained it.  The CTA-pair multicast loop (``gemm_persistent<2>``, validated on B200) runs through
the same checker as a control, and mutated protocols (a wrong arrival count, a missing remote arrive) must be caught.
"""
import random

import pytest


class ProtocolError(Exception):
    pass


class MBar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.tx, self.phase = name, count, count, 0, 0

    def _maybe_flip(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.count

    def arrive(self, expect_tx=0):
        if self.pending == 0:
            raise ProtocolError(f"{self.name}: arrival on a phase whose arrivals are already complete")
        self.tx += expect_tx
        self.pending -= 1
        self._maybe_flip()

    def complete_tx(self, nbytes):
        self.tx -= nbytes
        self._maybe_flip()

    def done(self, parity):  # mbarrier.try_wait.parity: has the phase with this parity completed?
        return (self.phase & 1) != parity


def run(agents, pending_events, rng, max_steps=200000):
    """agents: generators that yield a zero-argument predicate to wait on (or None to just yield the CPU).
    pending_events: list of callables (async completions) that the scheduler fires at random times, FIFO per queue."""
    waiting = {a: None for a in agents}
    steps = 0
    while waiting:
        steps += 1
        if steps > max_steps:
            raise ProtocolError("livelock")
        choices = [("agent", a) for a, cond in waiting.items() if cond is None or cond()]
        choices += [("event", q) for q in pending_events if q]
        if not choices:
            raise ProtocolError("deadlock: " + ", ".join(sorted(a.gi_code.co_name + str(a.gi_frame.f_locals.get("cta", ""))
                                                               for a in waiting)))
        kind, x = rng.choice(choices)
        if kind == "event":
            x.pop(0)()
            continue
        try:
            waiting[x] = next(x)
        except StopIteration:
            del waiting[x]
    for q in pending_events:
        while q:
            q.pop(0)()


A_BYTES, B_BYTES = 16384, 32768


def simulate_2sm(num_items, num_kb, stages, seed, full_count=2, peer_arrives=True, tmem_empty_count=8):
    rng = random.Random(seed)
    full = [MBar(f"full[{s}]", full_count) for s in range(stages)]                       # leader-owned
    empty = [[MBar(f"empty[{c}][{s}]", 1) for s in range(stages)] for c in range(2)]     # one set per CTA
    tmem_full = [[MBar(f"tmem_full[{c}][{a}]", 1) for a in range(2)] for c in range(2)]
    tmem_empty = [MBar(f"tmem_empty[{a}]", tmem_empty_count) for a in range(2)]           # leader-owned
    stage_state = [["free"] * stages for _ in range(2)]   # free -> loading -> (full barrier) -> reading -> free
    acc_state = ["free", "free"]                          # free -> accumulating -> full -> (drained by 8 warps) free
    acc_readers = [0, 0]
    tma_q = [[], []]          # per CTA: TMA completions in issue order
    tc_q = []                 # tensor-core commit arrivals, in issue order (tcgen05.commit tracks all prior MMAs)
    stage_bytes = A_BYTES + B_BYTES // 2

    def producer(cta):
        stage, phase = 0, 0
        for _ in range(num_items):
            for _ in range(num_kb):
                yield lambda s=stage, p=phase: empty[cta][s].done(p ^ 1)
                if stage_state[cta][stage] != "free":
                    raise ProtocolError(f"cta {cta}: stage {stage} refilled while {stage_state[cta][stage]}")
                stage_state[cta][stage] = "loading"
                if cta == 0:
                    full[stage].arrive(expect_tx=2 * stage_bytes)

                def landed(s=stage, c=cta):
                    stage_state[c][s] = "landed"
                    full[s].complete_tx(stage_bytes)

                tma_q[cta].append(landed)
                if cta == 1 and peer_arrives:
                    full[stage].arrive()
                stage += 1
                if stage == stages:
                    stage, phase = 0, phase ^ 1

    def mma():
        stage, phase = 0, 0
        for item in range(num_items):
            acc, acc_phase = item & 1, (item >> 1) & 1
            yield lambda a=acc, p=acc_phase: tmem_empty[a].done(p ^ 1)
            if acc_state[acc] != "free":
                raise ProtocolError(f"accumulator {acc} overwritten while {acc_state[acc]}")
            acc_state[acc] = "accumulating"
            for kb in range(num_kb):
                yield lambda s=stage, p=phase: full[s].done(p)
                for c in range(2):
                    if stage_state[c][stage] != "landed":
                        raise ProtocolError(f"MMA reads stage {stage} of cta {c} while {stage_state[c][stage]}")
                    stage_state[c][stage] = "reading"

                def stage_free(s=stage):
                    for c in range(2):
                        stage_state[c][s] = "free"
                        empty[c][s].arrive()

                tc_q.append(stage_free)
                if kb == num_kb - 1:
                    def acc_full(a=acc):
                        acc_state[a] = "full"
                        acc_readers[a] = 8
                        for c in range(2):
                            tmem_full[c][a].arrive()

                    tc_q.append(acc_full)
                stage += 1
                if stage == stages:
                    stage, phase = 0, phase ^ 1

    def epilogue(cta, warp):
        for item in range(num_items):
            acc, acc_phase = item & 1, (item >> 1) & 1
            yield lambda a=acc, p=acc_phase: tmem_full[cta][a].done(p)
            if acc_state[acc] != "full":
                raise ProtocolError(f"epilogue reads accumulator {acc} while {acc_state[acc]}")
            yield None  # tcgen05.ld + stores take a while
            acc_readers[acc] -= 1
            if acc_readers[acc] == 0:
                acc_state[acc] = "free"
            tmem_empty[acc].arrive()

    agents = [producer(0), producer(1), mma()] + [epilogue(c, w) for c in range(2) for w in range(4)]
    run(agents, tma_q + [tc_q], rng)
    return True


def simulate_pair_multicast(num_items, num_kb, stages, seed):
    """gemm_persistent<2>: every CTA runs its own MMAs on its own accumulator; B halves are multicast, so a stage of
    EITHER CTA may only be refilled when BOTH MMAs consumed it (empty count 2, commits multicast to the pair)."""
    rng = random.Random(seed)
    full = [[MBar(f"full[{c}][{s}]", 1) for s in range(stages)] for c in range(2)]
    empty = [[MBar(f"empty[{c}][{s}]", 2) for s in range(stages)] for c in range(2)]
    tmem_full = [[MBar(f"tmem_full[{c}][{a}]", 1) for a in range(2)] for c in range(2)]
    tmem_empty = [[MBar(f"tmem_empty[{c}][{a}]", 4) for a in range(2)] for c in range(2)]
    readers = [[0] * stages for _ in range(2)]      # MMAs that still have to consume stage s of cta c
    landed = [[0] * stages for _ in range(2)]
    tma_q, tc_q = [[], []], [[], []]

    def producer(cta):
        stage, phase = 0, 0
        for _ in range(num_items):
            for _ in range(num_kb):
                yield lambda s=stage, p=phase: empty[cta][s].done(p ^ 1)
                full[cta][stage].arrive(expect_tx=A_BYTES + B_BYTES)

                def own(s=stage, c=cta):
                    landed[c][s] += A_BYTES
                    full[c][s].complete_tx(A_BYTES)

                def half_b(s=stage):
                    for c in range(2):  # multicast: my half of B lands in both CTAs
                        if readers[c][s]:
                            raise ProtocolError(f"multicast into stage {s} of cta {c} while its MMA still reads it")
                        landed[c][s] += B_BYTES // 2
                        full[c][s].complete_tx(B_BYTES // 2)

                tma_q[cta] += [own, half_b]
                stage += 1
                if stage == stages:
                    stage, phase = 0, phase ^ 1

    def mma(cta):
        stage, phase = 0, 0
        for item in range(num_items):
            acc, acc_phase = item & 1, (item >> 1) & 1
            yield lambda a=acc, p=acc_phase: tmem_empty[cta][a].done(p ^ 1)
            for kb in range(num_kb):
                yield lambda s=stage, p=phase: full[cta][s].done(p)
                if landed[cta][stage] != A_BYTES + B_BYTES:
                    raise ProtocolError(f"cta {cta}: MMA on a stage with {landed[cta][stage]} bytes")
                readers[cta][stage] = 1

                def consumed(s=stage, c=cta):
                    readers[c][s] = 0
                    landed[c][s] = 0
                    for d in range(2):
                        empty[d][s].arrive()

                tc_q[cta].append(consumed)
                if kb == num_kb - 1:
                    tc_q[cta].append(lambda a=acc, c=cta: tmem_full[c][a].arrive())
                stage += 1
                if stage == stages:
                    stage, phase = 0, phase ^ 1

    def epilogue(cta, warp):
        for item in range(num_items):
            acc, acc_phase = item & 1, (item >> 1) & 1
            yield lambda a=acc, p=acc_phase: tmem_full[cta][a].done(p)
            yield None
            tmem_empty[cta][acc].arrive()

    agents = [producer(0), producer(1), mma(0), mma(1)] + [epilogue(c, w) for c in range(2) for w in range(4)]
    run(agents, tma_q + tc_q, rng)
    return True


@pytest.mark.parametrize("num_items,num_kb", [(1, 1), (1, 7), (2, 3), (3, 6), (5, 13), (4, 1)])
def test_two_sm_pipeline_has_no_deadlock_and_no_hazard(num_items, num_kb):
    for seed in range(25):
        assert simulate_2sm(num_items, num_kb, stages=6, seed=seed)


@pytest.mark.parametrize("num_items,num_kb", [(1, 1), (2, 5), (5, 9)])
def test_validated_pair_multicast_pipeline_passes_the_same_checker(num_items, num_kb):
    for seed in range(25):
        assert simulate_pair_multicast(num_items, num_kb, stages=4, seed=seed)


@pytest.mark.parametrize("mutation", [dict(full_count=1), dict(full_count=3), dict(peer_arrives=False),
                                      dict(tmem_empty_count=4), dict(tmem_empty_count=9)])
def test_checker_catches_broken_protocols(mutation):
    """Not vacuous: each of these one-line protocol errors is reported (as a deadlock or as a hazard)."""
    caught = 0
    for seed in range(40):
        try:
            simulate_2sm(4, 5, stages=6, seed=seed, **mutation)
        except ProtocolError:
            caught += 1
    assert caught > 0, mutation
