"""Multi-process host logic on CPU: gloo backend, world_size 2 and 4."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from baseline.ring_reference import ring_allreduce
    from hpc_patterns_b200.parallel.comm import Comm

    comm = Comm(backend="gloo")
    assert comm.world == world and comm.rank == rank
    comm.barrier()
    assert comm.max(float(rank)) == world - 1
    assert comm.min(float(rank) + 5) == 5
    assert comm.sum(1.0) == world
    objs = comm.all_gather_object({"rank": rank, "handle": bytes([rank]) * 64})
    assert [o["rank"] for o in objs] == list(range(world))
    assert objs[rank]["handle"] == bytes([rank]) * 64
    # the reference ring pattern: every element ends at P(P-1)/2 (float and int)
    for dtype in (torch.float32, torch.int32):
        va = torch.full((1000,), rank, dtype=dtype)
        vb = torch.full((1000,), rank, dtype=dtype)
        vc = torch.zeros(1000, dtype=dtype)
        ring_allreduce(va, vb, vc)
        assert bool((vc == world * (world - 1) // 2).all())
        vc.zero_()
        ring_allreduce(torch.full((1000,), rank, dtype=dtype), vb, vc, use_collective=True)
        assert bool((vc == world * (world - 1) // 2).all())
    comm.close()
    q.put((rank, "ok"))


@pytest.mark.parametrize("world,port", [(2, 29731), (4, 29732)])
def test_comm_and_ring_pattern_over_gloo(world, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert sorted(q.get(timeout=5)[0] for _ in range(world)) == list(range(world))


def test_single_process_comm():
    from hpc_patterns_b200.parallel.comm import Comm

    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        os.environ.pop(k, None)
    c = Comm()
    assert c.world == 1 and c.max(3.0) == 3.0 and c.all_gather_object("x") == ["x"]
    c.barrier()
