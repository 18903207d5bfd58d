"""Device timing helpers: CUDA events on the launching stream, max over ranks."""
from __future__ import annotations

from typing import Callable

import torch


def time_region_ms(fn: Callable[[], None], device: int, stream: torch.cuda.Stream | None = None) -> float:
    """Elapsed device milliseconds of everything ``fn`` enqueues on the current stream."""
    stream = stream or torch.cuda.current_stream(device)
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(device)
    e0.record(stream)
    fn()
    e1.record(stream)
    torch.cuda.synchronize(device)
    return float(e0.elapsed_time(e1))


def min_of(fn: Callable[[], float], iters: int, warmup: int = 3) -> float:
    for _ in range(warmup):
        fn()
    return min(fn() for _ in range(iters))


def flush_l2(device: int, nbytes: int = 256 << 20) -> None:
    """Evict L2 (126 MB on B200) by writing a larger scratch buffer."""
    buf = getattr(flush_l2, "_buf", None)
    if buf is None or buf.numel() < nbytes or buf.device.index != device:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=torch.device("cuda", device))
        flush_l2._buf = buf
    buf.fill_(1)
