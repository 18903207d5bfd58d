#!/usr/bin/env bash
# First 1-GPU call of the next round: validate what was written after the round-1 GPU budget ran out.
#   1. the opt-in 2-SM UMMA GEMM (tcgen05.mma.cta_group::2) under a short timeout (a protocol error = hang)
#   2. if it passes: time it next to the cluster-multicast kernel and cuBLAS, and take one ncu capture
#   3. the regular GPU suite + the N=1 bench (regression check)
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT
export HPCP_EXPERIMENTAL=1
if timeout 240 python -m pytest tests/test_gpu_kernels.py -k "2sm" -x -q --timeout 120 2>&1 | tail -5 | tee $OUT/next_2sm_pytest.txt | grep -q passed; then
  timeout 300 python scripts/gemm_put_bench.py 2>/dev/null | grep '^{' | tee $OUT/next_gemm_2sm.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.items() if 'tflops' in k or k in 'mnk'})"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_put_2sm -c 1 -f -o $OUT/prof_gemm_2sm \
    python -c "
import torch
from hpc_patterns_b200.ops.gemm import gemm_put
a = torch.randn(8192, 4096, device='cuda').bfloat16(); b = torch.randn(8192, 4096, device='cuda').bfloat16()
c = torch.empty(8192, 8192, device='cuda', dtype=torch.bfloat16)
for _ in range(3): gemm_put(a, b, c, 0, out_dtype=torch.bfloat16, cluster=3)
torch.cuda.synchronize()" > $OUT/next_ncu.log 2>&1
  ncu -i $OUT/prof_gemm_2sm.ncu-rep --page raw --csv > $OUT/prof_gemm_2sm.raw.csv 2>/dev/null
else
  echo "2-SM UMMA variant FAILED or hung: see $OUT/next_2sm_pytest.txt"
fi
# the plain GEMM->put as an instantiation of the policy template (same loop, different register allocation): if it
# passes the regular gemm_put tests the written-out copy in gemm_put.cu can be retired
HPCP_GEMM_PUT_TEMPLATE=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -k "gemm_put and not 2sm" -q --timeout 120 2>&1 | tail -3 | tee $OUT/next_gemm_put_template_pytest.txt
# L2 cache-policy variant of the flagship (opt-in): exactness, then N=1 timing next to the default
if timeout 120 python -m pytest tests/test_gpu_kernels.py -k "l2_hint" -x -q --timeout 60 2>&1 | tail -2 | tee $OUT/next_l2hint_pytest.txt | grep -q passed; then
  for r in 1 3; do
    timeout 200 python bench.py --gpus 1 --compute-ratio $r --no-extras --l2-hint 1 | tee $OUT/next_bench_n1_r${r}_l2hint.json | cut -c1-200
    timeout 200 python bench.py --gpus 1 --compute-ratio $r --no-extras | tee $OUT/next_bench_n1_r${r}_default.json | cut -c1-200
  done
fi
# GEMM fused with reduce-scatter / all-gather (written after the GPU budget ran out): one-GPU virtual-rank tests
# first (exactness of tile order, ownership, gather, counters), then a single-rank timing of both layers
timeout 600 python -m pytest tests/test_gpu_gemm_collective.py -q --timeout 120 2>&1 | tail -12 | tee $OUT/next_gemm_collective_pytest.txt
timeout 300 python scripts/tp_bench.py --check --mlp --tokens 8192 --out-features 8192 --in-features 4096 2>&1 | tail -2 | tee $OUT/next_tp_bench_n1.json | cut -c1-400
if grep -q passed $OUT/next_gemm_collective_pytest.txt; then   # ncu of the three fused kernels (virtual ranks, one GPU)
  for kern in gemm_reduce_scatter_kernel gemm_reduce_scatter_tma_kernel gemm_all_to_all_kernel allgather_gemm_kernel; do
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:$kern -c 1 -f -o $OUT/prof_$kern \
      python scripts/ncu_gemm_collective.py > $OUT/next_ncu_$kern.log 2>&1
    ncu -i $OUT/prof_$kern.ncu-rep --page raw --csv > $OUT/prof_$kern.raw.csv 2>/dev/null
  done
fi
unset HPCP_EXPERIMENTAL
# ncu of the validated GEMM kernel (CTA-pair multicast variant) for comparison with the 2-SM capture
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_put_kernel -c 1 -f -o $OUT/prof_gemm_put \
  python -c "
import torch
from hpc_patterns_b200.ops.gemm import gemm_put
a = torch.randn(8192, 4096, device='cuda').bfloat16(); b = torch.randn(8192, 4096, device='cuda').bfloat16()
c = torch.empty(8192, 8192, device='cuda', dtype=torch.bfloat16)
for _ in range(3): gemm_put(a, b, c, 0, out_dtype=torch.bfloat16)
torch.cuda.synchronize()" > $OUT/next_ncu_gemm.log 2>&1
ncu -i $OUT/prof_gemm_put.ncu-rep --page raw --csv > $OUT/prof_gemm_put.raw.csv 2>/dev/null
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6 | tee $OUT/next_pytest.txt
timeout 200 python bench.py --gpus 1 | tee $OUT/next_bench_n1.json | cut -c1-300
echo "== next-round check done"
