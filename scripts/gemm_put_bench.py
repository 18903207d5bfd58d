#!/usr/bin/env python
"""K-gemm-put vs the stock pattern (cuBLAS matmul, then a copy-engine peer copy).

1 GPU : GEMM-only throughput of the hand-written tcgen05 kernel next to torch.matmul (cuBLAS).
N GPUs: (torchrun) every rank computes C = A.B^T and puts it into its ring neighbour —
        fused in the kernel's epilogue vs matmul + cudaMemcpyAsync to the peer mapping.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hpc_patterns_b200.ops.gemm import gemm_put  # noqa: E402
from hpc_patterns_b200.parallel.comm import Comm  # noqa: E402
from hpc_patterns_b200.parallel.symmetric import SignalPads, SymmetricBuffer  # noqa: E402
from hpc_patterns_b200 import native  # noqa: E402


def timed(fn, comm, dev, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(dev)
    best = float("inf")
    for _ in range(iters):
        comm.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize(dev)
        best = min(best, comm.max(e0.elapsed_time(e1)))
    return best


def main():
    comm = Comm()
    dev = comm.local_rank % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    C = native()
    shapes = [(8192, 8192, 1024), (8192, 8192, 2048), (8192, 8192, 4096), (8192, 8192, 8192)]
    right, left = (comm.rank + 1) % comm.world, (comm.rank - 1) % comm.world
    pads = SignalPads(comm, dev)
    rows = []
    for m, n, k in shapes:
        a = torch.randn(m, k, device=f"cuda:{dev}").to(torch.bfloat16)
        b = torch.randn(n, k, device=f"cuda:{dev}").to(torch.bfloat16)
        c = torch.empty(m, n, device=f"cuda:{dev}")
        recv = SymmetricBuffer(comm, m * n * 4, dev, zero=False)
        flops = 2.0 * m * n * k
        t_ours = timed(lambda: gemm_put(a, b, c, 0), comm, dev)
        c_bf = torch.empty(m, n, device=f"cuda:{dev}", dtype=torch.bfloat16)
        t_cublas = timed(lambda: torch.matmul(a, b.t(), out=c_bf), comm, dev)
        row = {"m": m, "n": n, "k": k, "ranks": comm.world, "gemm_ms": t_ours, "gemm_tflops": flops / t_ours / 1e9,
               "cublas_bf16out_ms": t_cublas, "cublas_tflops": flops / t_cublas / 1e9}
        epoch = [0]

        def fused():
            epoch[0] += 1
            sync = pads.sync_ops(signal_rank=right, signal_section=C.PAD_DONE, epoch=epoch[0])
            pads.advance_tickets(gemm_put(a, b, None, recv.ptrs[right], sync=sync))
            C.wait(pads.word(comm.rank, C.PAD_DONE + left), epoch[0], pads.timeout_ns, pads.status_ptr,
                   torch.cuda.current_stream(dev).cuda_stream)

        c32 = torch.empty(m, n, device=f"cuda:{dev}")

        def stock():
            torch.matmul(a.float(), b.float().t(), out=c32) if False else None
            tmp = torch.matmul(a, b.t())                  # cuBLAS bf16 GEMM
            c32.copy_(tmp)                                # fp32 result like ours
            C.memcpy_async(recv.ptrs[right], c32.data_ptr(), m * n * 4, torch.cuda.current_stream(dev).cuda_stream)

        row["fused_gemm_put_ms"] = timed(fused, comm, dev)
        row["stock_cublas_then_memcpy_ms"] = timed(stock, comm, dev)
        row["speedup_vs_stock"] = row["stock_cublas_then_memcpy_ms"] / row["fused_gemm_put_ms"]
        pads.check()
        rows.append(row)
        if comm.rank == 0:
            print(json.dumps(row), flush=True)
        recv.close()
    pads.close()
    comm.close()


if __name__ == "__main__":
    main()
