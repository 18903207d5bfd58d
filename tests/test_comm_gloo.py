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


def test_comm_binds_the_device_the_mapping_chose(monkeypatch):
    """Dry run of the device choice (no GPU): under `tile_mapping ... SET` the wrapper exports HPCP_DEVICE, which differs
    from LOCAL_RANK for `spread` / `compact_plan` (8 GPUs, spread: rank 1 -> GPU 4).  Comm must pick THAT ordinal, wrap
    it to the visible devices, bind the process group to it (device_id) and make it current before NCCL is initialised."""
    import torch.distributed as dist

    from hpc_patterns_b200.parallel import comm as comm_mod
    from hpc_patterns_b200.parallel import tile_mapping as tm

    # the pure rule
    monkeypatch.setenv("HPCP_DEVICE", "4")
    assert comm_mod.Comm.pick_device(local_rank=1, n_devices=8) == 4
    assert comm_mod.Comm.pick_device(local_rank=1, n_devices=2) == 0            # wrapped to what is visible (CVD)
    monkeypatch.delenv("HPCP_DEVICE")
    assert comm_mod.Comm.pick_device(local_rank=5, n_devices=8) == 5
    assert comm_mod.Comm.pick_device(local_rank=9, n_devices=8) == 1            # more ranks than GPUs
    assert comm_mod.Comm.pick_device(local_rank=3, n_devices=0) == 0            # CPU box

    # the wiring, with a box of 8 fake GPUs
    calls = {}
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "device_count", lambda: 8)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: calls.setdefault("set_device", d))
    monkeypatch.setattr(dist, "is_initialized", lambda: False)
    monkeypatch.setattr(dist, "init_process_group", lambda backend, **kw: calls.update(backend=backend, **kw))
    monkeypatch.setattr(dist, "get_backend", lambda: "nccl")
    for k, v in (("RANK", "1"), ("WORLD_SIZE", "8"), ("LOCAL_RANK", "1")):
        monkeypatch.setenv(k, v)
    monkeypatch.setenv("HPCP_DEVICE", str(tm.device_for_rank("spread", 1, 8)))
    c = comm_mod.Comm()
    assert c.device == 4 != c.local_rank
    assert calls["set_device"] == 4 and calls["device_id"] == torch.device("cuda", 4)
    assert calls["rank"] == 1 and calls["world_size"] == 8 and "nccl" in calls["backend"]
    # an explicit device wins over the environment
    calls.clear()
    assert comm_mod.Comm(device=6).device == 6 and calls["device_id"] == torch.device("cuda", 6)
