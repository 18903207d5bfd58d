"""K-p2p: in-kernel GPU<->GPU copies (put / get) over NVLink peer mappings.

Replaces ``MPI_Put``/``MPI_Isend`` on device pointers in the reference
(p2p/peer2pear.cpp:32-44,76-81).  ``dst``/``src`` are tensors or raw addresses;
either may be a peer-mapped address obtained from ``parallel.symmetric``.
"""
from __future__ import annotations

from typing import Optional

import torch

from .. import native
from ._util import PtrLike, current_stream, ptr


def copy(dst: PtrLike, src: PtrLike, nbytes: int, *, src_is_peer: bool = False, engine: str = "ldst",
         tune: Optional[dict] = None, sync: Optional[dict] = None, device: int = 0,
         stream: Optional[int] = None) -> int:
    """dst[0:nbytes] = src[0:nbytes]; returns the number of CTAs launched."""
    C = native()
    return C.copy(ptr(dst), ptr(src), int(nbytes), src_is_peer, engine, tune or {}, sync or {}, device,
                  current_stream(device) if stream is None else stream)


def fill_pattern(dst: PtrLike, n_words: int, seed: int, device: int = 0, stream: Optional[int] = None) -> None:
    native().fill_pattern(ptr(dst), int(n_words), seed & 0xFFFFFFFF,
                          current_stream(device) if stream is None else stream)


def pattern_reference(n_words: int, seed: int) -> torch.Tensor:
    """Plain PyTorch (CPU, int64 arithmetic) reference of the payload pattern."""
    i = torch.arange(n_words, dtype=torch.int64)
    m = 0xFFFFFFFF
    x = ((i * 2654435761) & m) ^ (seed & m)
    x = x ^ (x >> 16)
    x = (x * 0x7FEB352D) & m
    x = x ^ (x >> 15)
    x = (x * 0x846CA68B) & m
    x = x ^ (x >> 16)
    return x


def verify_pattern(data: PtrLike, n_words: int, seed: int, device: int = 0,
                   stream: Optional[int] = None) -> int:
    """Exact device-side check; returns the number of mismatching 32-bit words."""
    C = native()
    counters = torch.zeros(2, dtype=torch.int64, device=torch.device("cuda", device))
    C.verify_pattern(ptr(data), int(n_words), seed & 0xFFFFFFFF, counters.data_ptr(),
                     counters.data_ptr() + 8, 0, 0, 0, 0,
                     current_stream(device) if stream is None else stream)
    return int(counters[0].item())
