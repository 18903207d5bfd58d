#!/usr/bin/env python
"""Tensor-parallel linear layers: fused kernels vs the stock pattern, N ranks (torchrun) or 1 GPU.

  row-parallel     y = reduce_scatter(x_r @ W_r.T)   K-gemm-rs  vs  cuBLAS matmul + NCCL reduce_scatter
                   y = all_reduce(x_r @ W_r.T)       K-gemm-ar (multimem.red through the switch)  vs  + NCCL all_reduce
  column-parallel  y = all_gather(x_r) @ W_r.T       K-ag-gemm  vs  NCCL all_gather + cuBLAS matmul

`--check` compares both fused layers with the stock result on exactly representable operands (exact equality),
then times them: device events, max over ranks, min over iterations.  One JSON line from rank 0.

  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/tp_bench.py --m 8192 --n 8192 --k 28672
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hpc_patterns_b200.models.tensor_parallel import ColumnParallelLinear, RowParallelLinear  # noqa: E402
from hpc_patterns_b200.parallel.comm import Comm  # noqa: E402


def timed(fn, comm, dev, iters, warm=3):
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


def dyadic(shape, dev, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    return (torch.randint(-4, 5, shape, device=dev, generator=g).float() / 4).to(torch.bfloat16)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=8192, help="rows (tokens) of the whole problem")
    ap.add_argument("--n", type=int, default=8192, help="output features of the whole problem")
    ap.add_argument("--k", type=int, default=8192, help="input features of the whole problem")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--cluster", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=0, help="all-gather granularity in bytes (0 -> 4096)")
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()

    comm = Comm()
    dev = comm.local_rank % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    device = torch.device("cuda", dev)
    P = comm.world
    out = {"ranks": P, "m": args.m, "n": args.n, "k": args.k}

    # ---- row-parallel: K is sharded, the output rows are scattered -------------------------------
    k_local = args.k // P
    row = RowParallelLinear(comm, dev, args.m, args.n, k_local, cluster=args.cluster)
    x = dyadic((args.m, k_local), device, 100 + comm.rank)
    row.w.copy_(dyadic((args.n, k_local), device, 200 + comm.rank))
    if args.check:
        y = row.forward(x).clone()
        ref = row.stock_forward(x)
        if P == 1:
            ref = ref[: args.m // P]
        torch.cuda.synchronize(dev)
        row.check()
        out["row_parallel_exact"] = bool(comm.min(float(torch.equal(y, ref))) == 1.0)
    t_fused = timed(lambda: row.forward(x), comm, dev, args.steps)
    ref_out = torch.empty(args.m // P, args.n, device=device)
    t_stock = timed(lambda: row.stock_forward(x, ref_out), comm, dev, args.steps)
    row.check()
    flops = 2.0 * args.m * args.n * k_local
    out["row_parallel"] = {"fused_ms": round(t_fused, 4), "stock_ms": round(t_stock, 4),
                           "speedup": round(t_stock / t_fused, 3),
                           "fused_tflops_per_gpu": round(flops / t_fused / 1e9, 1),
                           "nvlink_GBps_per_gpu": round(args.m * args.n * 4 * (P - 1) / P / (t_fused * 1e6), 1)}
    row.close()

    # ---- row-parallel with an all-reduce through the switch (NVLS) -------------------------------
    if P > 1:
        try:
            ar = RowParallelLinear(comm, dev, args.m, args.n, k_local, cluster=args.cluster, reduce="all")
        except RuntimeError as e:  # no multicast support on this system
            ar = None
            out["row_parallel_allreduce"] = {"unavailable": str(e)[:120]}
        if ar is not None:
            ar.w.copy_(dyadic((args.n, k_local), device, 200 + comm.rank))
            if args.check:
                y = ar.forward(x).clone()
                ref = ar.stock_forward(x)
                torch.cuda.synchronize(dev)
                ar.check()
                out["row_parallel_allreduce_exact"] = bool(comm.min(float(torch.equal(y, ref))) == 1.0)
            t_fused = timed(lambda: ar.forward(x), comm, dev, args.steps)
            t_stock = timed(lambda: ar.stock_forward(x), comm, dev, args.steps)
            ar.check()
            out["row_parallel_allreduce"] = {"fused_ms": round(t_fused, 4), "stock_ms": round(t_stock, 4),
                                             "speedup": round(t_stock / t_fused, 3),
                                             "nvlink_out_GBps_per_gpu": round(args.m * args.n * 4 / (t_fused * 1e6), 1)}
            ar.close()

    # ---- column-parallel: rows of x are sharded and gathered, N is sharded -------------------------
    n_local = args.n // P
    col = ColumnParallelLinear(comm, dev, args.m, n_local, args.k, out_dtype=torch.bfloat16, cluster=args.cluster,
                               chunk_bytes=args.chunk)
    x_rows = dyadic((args.m // P, args.k), device, 300 + comm.rank)
    col.w.copy_(dyadic((n_local, args.k), device, 400 + comm.rank))
    if args.check:
        y = col.forward(x_rows).clone()
        ref = col.stock_forward(x_rows)
        torch.cuda.synchronize(dev)
        col.check()
        out["column_parallel_exact"] = bool(comm.min(float(torch.equal(y, ref.to(y.dtype)))) == 1.0)
    t_fused = timed(lambda: col.forward(x_rows), comm, dev, args.steps)
    t_stock = timed(lambda: col.stock_forward(x_rows), comm, dev, args.steps)
    col.check()
    flops = 2.0 * args.m * n_local * args.k
    out["column_parallel"] = {"fused_ms": round(t_fused, 4), "stock_ms": round(t_stock, 4),
                              "speedup": round(t_stock / t_fused, 3),
                              "fused_tflops_per_gpu": round(flops / t_fused / 1e9, 1),
                              "nvlink_GBps_per_gpu": round(args.m * args.k * 2 * (P - 1) / P / (t_fused * 1e6), 1)}
    col.close()
    if comm.rank == 0:
        print(json.dumps(out), flush=True)
    ok = (not args.check) or (out["row_parallel_exact"] and out["column_parallel_exact"] and
                              out.get("row_parallel_allreduce_exact", True))
    comm.close()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
