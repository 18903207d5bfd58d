#!/usr/bin/env python
"""Tensor-parallel linear layers: fused kernels vs the stock pattern, N ranks (torchrun) or 1 GPU.

  row-parallel     y = reduce_scatter(x_r @ W_r.T)   K-gemm-rs  vs  cuBLAS matmul + NCCL reduce_scatter
                   y = all_reduce(x_r @ W_r.T)       K-gemm-ar (multimem.red through the switch)  vs  + NCCL all_reduce
  column-parallel  y = all_gather(x_r) @ W_r.T       K-ag-gemm  vs  NCCL all_gather + cuBLAS matmul

`--check` compares the fused layers with the stock result on exactly representable operands (exact equality),
then times them: device events, max over ranks, min over iterations.  One JSON line from rank 0.
Same as `python -m hpc_patterns_b200 tp ...`.

  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/tp_bench.py --tokens 8192 --out-features 8192 --in-features 28672
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hpc_patterns_b200.models.tensor_parallel import main  # noqa: E402

if __name__ == "__main__":
    sys.exit(main())
