"""K-gemm-put: tcgen05 bf16 GEMM whose epilogue stores the result tile into a peer GPU.

``gemm_put(a, b, c_local=..., c_peer=...)`` computes ``C = A @ B.T`` (A ``[M,K]``, B ``[N,K]``, both
bf16 row-major, fp32 result) with the hand-written sm_100a kernel in ``csrc/kernels/gemm_put.cu``
(TMA-fed smem ring, ``tcgen05.mma`` accumulating in TMEM, ``tcgen05.ld`` epilogue) and writes every
128x256 tile to ``c_local`` and/or straight into the peer-mapped ``c_peer`` over NVLink.
The reference has no GEMM; this op is the tensor-core member of the suite's fused
"produce a tile -> put it to the neighbour" family.  Stock comparison: ``torch.matmul`` (cuBLAS)
followed by a copy-engine peer copy.
"""
from __future__ import annotations

from typing import Optional

import torch

from .. import native
from ._util import PtrLike, current_stream, ptr


def gemm_put(a: torch.Tensor, b: torch.Tensor, c_local: Optional[torch.Tensor] = None, c_peer: PtrLike = 0,
             sync: Optional[dict] = None, ctas: int = 0, stream: Optional[int] = None,
             out_dtype: torch.dtype = torch.float32, cluster: int = 0) -> int:
    """Launch the fused GEMM(+put).  ``out_dtype`` fp32 or bf16 (c_local / c_peer hold that type).
    ``cluster``: 0 auto, 1 = single CTAs, 2 = CTA pairs sharing the B tile through TMA multicast,
    3 = 2-SM UMMA (``tcgen05.mma.cta_group::2``: one 256x256 tile per CTA pair; opt-in).
    Returns the number of CTAs launched (for ticket bookkeeping)."""
    if out_dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("out_dtype must be float32 or bfloat16")
    if a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16:
        raise TypeError("gemm_put takes bf16 operands")
    if a.dim() != 2 or b.dim() != 2 or a.shape[1] != b.shape[1]:
        raise ValueError("expected A[M,K] and B[N,K]")
    m, k = a.shape
    n = b.shape[0]
    if m % 128 or n % 256 or k % 64:
        raise ValueError("M, N, K must be multiples of 128, 256, 64")
    if c_local is not None and (c_local.dtype != out_dtype or tuple(c_local.shape) != (m, n)):
        raise ValueError("c_local must be [M,N] of out_dtype")
    dev = a.device.index
    return native().gemm_put(ptr(a), ptr(b), ptr(c_local) if c_local is not None else 0,
                             ptr(c_peer) if not isinstance(c_peer, int) else c_peer, m, n, k,
                             out_dtype == torch.bfloat16, sync or {}, ctas, dev,
                             current_stream(dev) if stream is None else stream, cluster)


def gemm_reference(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Plain PyTorch fp32 reference of the op."""
    return a.float() @ b.float().t()
