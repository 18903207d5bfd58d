from __future__ import annotations

from typing import Union

import torch

PtrLike = Union[int, torch.Tensor]


def ptr(x: PtrLike) -> int:
    if isinstance(x, torch.Tensor):
        if not x.is_contiguous():
            raise ValueError("native kernels need contiguous tensors")
        return x.data_ptr()
    return int(x)


def current_stream(device: int) -> int:
    return torch.cuda.current_stream(device).cuda_stream
