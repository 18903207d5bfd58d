"""The reference allreduce *pattern* through stock torch.distributed calls (gloo on CPU, NCCL on GPU).

This is the like-for-like baseline for K-ring: accumulate; then P-1 times
{send block right / receive block from left (odd ranks send first, even ranks receive first,
allreduce-mpi-sycl.cpp:43-59); swap; accumulate}, with a blocking wait after every step — or
``all_reduce`` when ``use_collective`` (the ``-a`` path, :61-67).  It runs on CPU tensors too,
which is how the host-side step logic is tested without a GPU.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def ring_allreduce(va: torch.Tensor, vb: torch.Tensor, vc: torch.Tensor, use_collective: bool = False) -> torch.Tensor:
    """In-place on vc; returns vc.  va/vb/vc are this rank's blocks (same shape/dtype)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    if use_collective:
        vc.copy_(va)
        dist.all_reduce(vc)
        return vc
    right, left = (rank + 1) % world, (rank - 1) % world
    vc.add_(va)
    for _ in range(1, world):
        if rank % 2:
            dist.send(va, right)
            dist.recv(vb, left)
        else:
            dist.recv(vb, left)
            dist.send(va, right)
        va, vb = vb, va
        vc.add_(va)
        if vc.is_cuda:
            torch.cuda.synchronize()
    return vc
