#!/usr/bin/env bash
# 1-GPU call: the 2-SM UMMA loop after the barrier fix (one arrival per full barrier, cta-scope waits) — exactness, then
# throughput next to the 1-SM kernel and cuBLAS; the new stencil_step op test.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/../..}"
OUT=gpurun_out/r2c7; mkdir -p $OUT
export PYTHONPATH=$PWD
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_gemm_collective.py tests/test_gpu_halo.py -q --timeout 120 -k "gemm or 2sm or stencil_step or allgather" > $OUT/pytest_gemm.txt 2>&1; tail -5 $OUT/pytest_gemm.txt
timeout 300 python scripts/gemm_put_bench.py 2>/dev/null | grep '^{' | tee $OUT/gemm_put_bench.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: (round(v, 1) if isinstance(v, float) else v) for k, v in d.items() if 'tflops' in k or k in ('k',) or 'diff' in k})"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_put_2sm -s 1 -c 1 -f -o $OUT/prof_gemm_2sm_fixed python gpurun_out/r2c6/gemm_one.py 3 > $OUT/ncu.log 2>&1 || true
cat > $OUT/gemm_one.py <<'PY'
import sys, torch
from hpc_patterns_b200.ops.gemm import gemm_put
cluster = int(sys.argv[1])
a = torch.randn(8192, 4096, device='cuda').bfloat16(); b = torch.randn(8192, 4096, device='cuda').bfloat16()
c = torch.empty(8192, 8192, device='cuda', dtype=torch.bfloat16)
for _ in range(3): gemm_put(a, b, c, 0, out_dtype=torch.bfloat16, cluster=cluster)
torch.cuda.synchronize()
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_put_2sm -s 1 -c 1 -f -o $OUT/prof_gemm_2sm_fixed python $OUT/gemm_one.py 3 > $OUT/ncu.log 2>&1; tail -1 $OUT/ncu.log
echo "== r2 call7 done"
