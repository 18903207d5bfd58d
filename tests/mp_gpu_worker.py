"""torchrun worker: process-per-GPU paths (CUDA IPC peer memory + in-kernel signalling)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from hpc_patterns_b200.models.allreduce import AllreduceMiniapp  # noqa: E402
from hpc_patterns_b200.models.halo import HaloStencil  # noqa: E402
from hpc_patterns_b200.models.peer2pear import FusedTriadExchange, P2PBench  # noqa: E402
from hpc_patterns_b200.parallel.comm import Comm  # noqa: E402


def main():
    comm = Comm()
    dev = comm.device
    torch.cuda.set_device(dev)
    # flagship: stencil fused with the halo exchange, across processes (CUDA IPC peer mappings) ---------------
    for mode in ("pull", "push"):
        for rows, nbytes, tune in ((1, 1 << 20, {"tile_kb": 4, "stages": 6}), (3, (8 << 20) + 4096, {})):
            hs = HaloStencil(comm, dev, message_bytes=nbytes, rows=rows, mode=mode, tune=tune)
            hs.step(1)
            hs.step(6)                       # persistent: six steps, one launch, neighbours spin on each other
            hs.step(1)
            torch.cuda.synchronize()
            hs.check()
            assert hs.verify_from_init() == 0, (mode, rows)
            comm.barrier()
            assert hs.verify_last_step() == 0, (mode, rows)
            comm.barrier()
            hs.reset()
            for how in ("memcpy", "nccl"):
                hs.stock_step(how)
            torch.cuda.synchronize()
            assert hs.verify_from_init() == 0, (mode, rows, "stock")
            comm.barrier()
            assert hs.verify_last_step() == 0, (mode, rows, "stock")
            comm.barrier()
            if mode == "push":
                hs.reset()
                bufs = hs.make_host_buffers()
                for i in range(3):
                    hs.step_from_host(bufs[i & 1], bufs[(i + 1) & 1], chunks=4)
                assert hs.verify_from_init() == 0
                assert torch.equal(bufs[1], hs.u_tensor().cpu())
            hs.close()
    if comm.rank == 0:
        print("HALO OK")
    # peer2pear, every transport
    for transport in ("put", "get", "sendrecv", "memcpy", "nccl"):
        for engine in (("ldst", "tma") if transport in ("put", "get", "sendrecv") else ("ldst",)):
            b = P2PBench(comm, dev, max_bytes=8 << 20, transport=transport, engine=engine, iters=3)
            for nbytes in (1024, 8 << 20):
                r = b.run(nbytes)
                assert r.mismatches == 0, (transport, engine, nbytes, r.mismatches)
                assert r.uni_gbps > 0 and r.bi_gbps > 0
            b.close()
    # fused exchange over NVLink
    for engine, ratio in (("ldst", 1), ("tma", 1), ("ldst", 3), ("tma", 3)):
        ex = FusedTriadExchange(comm, dev, nbytes=16 << 20, engine=engine, compute_ratio=ratio)
        for _ in range(4):
            ex.step()
        torch.cuda.synchronize()
        ex.check()
        assert ex.verify() == 0
        assert ex.step_from_host(ex.make_host_input(), chunks=4) == 0
        ex.put_only(); ex.triad_only(); ex.stock_step("memcpy"); ex.stock_step("nccl")
        torch.cuda.synchronize()
        ex.check()
        ex.close()
    # allreduce miniapp
    algos = ["ring", "ring-unfused", "twoshot", "nccl", "ring-nccl"]
    for algo in algos:
        for dtype in ("float", "int"):
            app = AllreduceMiniapp(comm, dev, log2_elems=18, dtype=dtype, algo=algo)
            res = app.run(iters=2, warmup=1)
            assert res.mismatches == 0, (algo, dtype)
            app.close()
    try:
        app = AllreduceMiniapp(comm, dev, log2_elems=18, dtype="float", algo="nvls")
        res = app.run(iters=2, warmup=1)
        assert res.mismatches == 0
        if comm.rank == 0:
            print("NVLS OK", res.ms)
        app.close()
    except Exception as e:  # multicast is optional on a given box
        if comm.rank == 0:
            print("NVLS unavailable:", repr(e)[:300])
    comm.barrier()
    if comm.rank == 0:
        print("WORKER OK")
    comm.close()


if __name__ == "__main__":
    main()
