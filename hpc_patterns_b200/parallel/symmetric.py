"""Symmetric (peer-mapped) buffers and signal pads for one-process-per-GPU runs.

Every rank allocates the same number of bytes with the native allocator
(``cudaMalloc``), exports a CUDA IPC handle, all-gathers the handles over the
host process group and opens its peers' handles: afterwards ``ptrs[r]`` is an
address this rank's kernels can load from / store to over NVLink.  This is the
one-process-per-GPU twin of ``NodeMemory`` in csrc/common/peer_mem.h and plays
the role GPU-aware MPICH's IPC layer plays for the reference
(``MPI_Win_create`` on a device buffer, p2p/peer2pear.cpp:119-122).
"""
from __future__ import annotations

from typing import List, Optional

import torch

from .. import native
from .comm import Comm


class _CudaArrayView:
    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {
            "shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}


def tensor_from_ptr(ptr: int, nbytes: int, device: int, dtype: torch.dtype = torch.uint8) -> torch.Tensor:
    """Zero-copy torch view of raw device memory (the memory must outlive the tensor)."""
    t = torch.as_tensor(_CudaArrayView(ptr, nbytes), device=torch.device("cuda", device))
    return t.view(dtype)


class SymmetricBuffer:
    """``nbytes`` on every rank; ``ptrs[r]`` = address of rank r's copy, usable from this rank."""

    def __init__(self, comm: Comm, nbytes: int, device: int, zero: bool = True):
        self.C = native()
        self.comm = comm
        self.nbytes = int(nbytes)
        self.device = device
        self.rank, self.world = comm.rank, comm.world
        self.local_ptr: int = self.C.alloc(self.nbytes, "D", device, zero)
        self._opened: List[int] = []
        if self.world == 1:
            self.ptrs = [self.local_ptr]
        else:
            handles = comm.all_gather_object(self.C.ipc_export(self.local_ptr))
            self.ptrs = []
            for r, h in enumerate(handles):
                if r == self.rank:
                    self.ptrs.append(self.local_ptr)
                else:
                    p = self.C.ipc_open(h)
                    self._opened.append(p)
                    self.ptrs.append(p)
            comm.barrier()

    @classmethod
    def local_group(cls, comms: list, nbytes: int, devices: List[int], zero: bool = True) -> List["SymmetricBuffer"]:
        """In-process twin for ``parallel.local.LocalGroup``: one allocation per virtual rank (several may share a
        GPU), every view sees all of them through plain pointers (peer access must be enabled between the GPUs)."""
        C = native()
        local = [C.alloc(int(nbytes), "D", dev, zero) for dev in devices]
        out = []
        for comm, dev, ptr in zip(comms, devices, local):
            b = cls.__new__(cls)
            b.C, b.comm, b.nbytes, b.device = C, comm, int(nbytes), dev
            b.rank, b.world = comm.rank, comm.world
            b.local_ptr, b._opened, b.ptrs = ptr, [], list(local)
            out.append(b)
        return out

    def tensor(self, dtype: torch.dtype = torch.uint8) -> torch.Tensor:
        return tensor_from_ptr(self.local_ptr, self.nbytes, self.device, dtype)

    def close(self) -> None:
        if self.local_ptr:
            torch.cuda.synchronize(self.device)
            self.comm.barrier()
            for p in self._opened:
                self.C.ipc_close(p)
            self._opened = []
            self.comm.barrier()
            self.C.free(self.local_ptr, "D")
            self.local_ptr = 0


class SignalPads:
    """Per-rank signal pad (see csrc/common/signal.cuh) + status word + epoch bookkeeping."""

    def __init__(self, comm: Comm, device: int, extra_words: int = 0, timeout_s: float = 20.0,
                 buf: Optional[SymmetricBuffer] = None):
        self.C = native()
        self.comm = comm
        self.device = device
        self.rank, self.world = comm.rank, comm.world
        self.extra_words = int(extra_words)
        self.buf = buf if buf is not None else SymmetricBuffer(comm, self.pad_bytes(extra_words), device, zero=True)
        self.timeout_ns = int(timeout_s * 1e9)
        self.ticket_issued = 0
        self.barrier_epoch = 0

    @staticmethod
    def pad_bytes(extra_words: int = 0) -> int:
        C = native()
        return 4 * (C.PAD_WORDS + int(extra_words) + C.PAD_TAIL_WORDS)

    def word(self, rank: int, index: int) -> int:
        return self.buf.ptrs[rank] + 4 * index

    @property
    def status_ptr(self) -> int:
        return self.word(self.rank, self.C.PAD_WORDS + self.extra_words)

    @property
    def ticket_ptr(self) -> int:
        return self.word(self.rank, self.C.PAD_LOCAL)

    def chunk_word(self, rank: int, chunk: int = 0) -> int:
        return self.word(rank, self.C.PAD_WORDS + chunk)

    def sync_ops(self, *, signal_rank: Optional[int] = None, signal_section: Optional[int] = None,
                 epoch: int = 0, wait_section: Optional[int] = None, wait_rank: Optional[int] = None) -> dict:
        """Build the kwargs dict understood by the native launchers."""
        d = {"timeout_ns": self.timeout_ns, "status": self.status_ptr,
             "ticket": self.ticket_ptr, "ticket_base": self.ticket_issued & 0xFFFFFFFF}
        if signal_rank is not None:
            d["signal_flag"] = self.word(signal_rank, signal_section + self.rank)
            d["signal_epoch"] = epoch
        if wait_section is not None:
            d["wait_flag"] = self.word(self.rank, wait_section + wait_rank)
            d["wait_epoch"] = epoch
        return d

    def advance_tickets(self, ctas: int) -> None:
        self.ticket_issued += int(ctas)

    def device_barrier(self, stream: int) -> None:
        """In-kernel barrier across all GPUs, enqueued on ``stream``."""
        self.barrier_epoch += 1
        self.C.barrier_all([self.buf.ptrs[r] for r in range(self.world)], self.rank,
                           self.barrier_epoch, self.timeout_ns, self.status_ptr, stream)

    def check(self) -> None:
        st = self.C.read_u32(self.status_ptr)
        if st != self.C.STATUS_OK:
            kind = "timeout waiting for a peer" if st == self.C.STATUS_TIMEOUT else f"status 0x{st:08x}"
            raise RuntimeError(f"rank {self.rank}: device-side synchronisation failed ({kind})")

    def close(self) -> None:
        self.buf.close()
