#!/usr/bin/env python
"""One launch of each fused collective GEMM on ONE GPU with two virtual ranks (peer pointers are local buffers), so
that `ncu -k regex:...` can capture the kernels without NVLink:  reduce-scatter (REDG epilogue), all-to-all (store
epilogue with per-tile destination) and all-gather (gather thread active: rank 0 pulls "rank 1's" rows).
Not a benchmark — numbers taken under a profiler are never reported."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hpc_patterns_b200.ops.gemm import allgather_gemm, gemm_all_to_all, gemm_reduce_scatter  # noqa: E402

m, n, k, world = 8192, 2048, 4096, 2
dev = torch.device("cuda", 0)
cluster = int(os.environ.get("HPCP_NCU_CLUSTER", "0"))
a = torch.randn(m, k, device=dev).bfloat16()
b = torch.randn(n, k, device=dev).bfloat16()
shards = [torch.zeros(m // world, n, device=dev) for _ in range(world)]
recv = [torch.zeros(world, m // world, n, device=dev, dtype=torch.bfloat16) for _ in range(world)]
a_other = torch.randn(m // world, k, device=dev).bfloat16()          # "rank 1's" row block
a_full = torch.zeros(m, k, device=dev, dtype=torch.bfloat16)
a_full[: m // world] = a[: m // world]
c = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
ready = torch.zeros(m // 128, dtype=torch.int32, device=dev)
per_launch = 128 * k * 2 // 4096
for it in range(3):
    gemm_reduce_scatter(a, b, shards, 0, cluster=cluster)
    gemm_reduce_scatter(a, b, shards, 0, cluster=cluster, epilogue="tma")
    gemm_all_to_all(a, b, recv, 0, out_dtype=torch.bfloat16, cluster=cluster)
    allgather_gemm(a_full, [a_full[: m // world], a_other], b, c, 0, ready=ready, ready_base=it * per_launch,
                   timeout_ns=int(5e9), cluster=cluster, activation="gelu")
torch.cuda.synchronize()
print("ok")
