"""One-process-per-GPU plumbing on top of ``torch.distributed``.

The reference bootstraps with ``MPI_Init/Comm_rank/Comm_size`` and uses MPI only
for barriers and host-scalar reductions of timings in its drivers
(p2p/peer2pear.cpp:26,49-50,107-110; allreduce-mpi-sycl.cpp:91-93,189).  Here the
same four things (rank, size, barrier, min/max/sum of a host scalar, plus an
object all-gather used to exchange CUDA IPC handles) come from a process group:
``nccl`` + ``gloo`` on GPU boxes, ``gloo`` alone on CPU (tests).  No payload data
ever travels through this layer — kernels move the data over NVLink themselves.
"""
from __future__ import annotations

import datetime
import os
from typing import Any, List

import torch
import torch.distributed as dist


class Comm:
    """rank / world / barrier / scalar reductions / object all-gather."""

    def __init__(self, backend: str | None = None, timeout_s: float = 600.0, device: int | None = None):
        """``device``: the CUDA ordinal this rank computes on.  Default: what the rank->device wrapper chose
        (``HPCP_DEVICE`` under ``tile_mapping ... SET``), else LOCAL_RANK, wrapped to the visible devices — the
        process group is bound to THAT device, so NCCL collectives and the kernels agree on the GPU
        (``spread`` on 8 GPUs puts rank 1 on GPU 4, not on GPU LOCAL_RANK)."""
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", str(self.rank)))
        self._owns_group = False
        self.backend = "single"
        self.device = self.pick_device(self.local_rank) if device is None else int(device)
        if self.world > 1:
            if not dist.is_initialized():
                if backend is None:
                    backend = "cpu:gloo,cuda:nccl" if torch.cuda.is_available() else "gloo"
                os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
                os.environ.setdefault("MASTER_PORT", "29511")
                kwargs = {}
                if torch.cuda.is_available():
                    torch.cuda.set_device(self.device)
                    kwargs["device_id"] = torch.device("cuda", self.device)
                dist.init_process_group(backend, rank=self.rank, world_size=self.world,
                                        timeout=datetime.timedelta(seconds=timeout_s), **kwargs)
                self._owns_group = True
            self.backend = dist.get_backend()

    @staticmethod
    def pick_device(local_rank: int, n_devices: int | None = None) -> int:
        from .tile_mapping import selected_device

        if n_devices is None:
            n_devices = torch.cuda.device_count() if torch.cuda.is_available() else 0
        return selected_device(default=local_rank) % max(n_devices, 1)

    # -- collectives on host scalars / objects -------------------------------------------
    def barrier(self) -> None:
        if self.world > 1:
            t = torch.zeros(1)
            dist.all_reduce(t)  # CPU tensor -> gloo path; no GPU work enqueued

    def _reduce(self, value: float, op) -> float:
        if self.world == 1:
            return float(value)
        t = torch.tensor([float(value)], dtype=torch.float64)
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, value: float) -> float:
        return self._reduce(value, dist.ReduceOp.MAX)

    def min(self, value: float) -> float:
        return self._reduce(value, dist.ReduceOp.MIN)

    def sum(self, value: float) -> float:
        return self._reduce(value, dist.ReduceOp.SUM)

    def all_gather_object(self, obj: Any) -> List[Any]:
        if self.world == 1:
            return [obj]
        out: List[Any] = [None] * self.world
        dist.all_gather_object(out, obj)
        return out

    def close(self) -> None:
        if self._owns_group and dist.is_initialized():
            dist.destroy_process_group()
            self._owns_group = False
