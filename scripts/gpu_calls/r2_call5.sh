#!/usr/bin/env bash
# Round 2, 1-GPU call after the eager-module-loading fix: the whole suite (virtual-rank tests at full width), ncu of the
# two GEMM tile loops (1-SM CTA-pair multicast vs 2-SM UMMA) and of the fused concurrency kernel.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/../..}"
OUT=gpurun_out/r2c5; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > $OUT/pytest_gpu_full.txt 2>&1; tail -12 $OUT/pytest_gpu_full.txt
cat > /tmp/gemm_one.py <<'PY'
import sys, torch
from hpc_patterns_b200.ops.gemm import gemm_put
cluster = int(sys.argv[1])
a = torch.randn(8192, 4096, device='cuda').bfloat16(); b = torch.randn(8192, 4096, device='cuda').bfloat16()
c = torch.empty(8192, 8192, device='cuda', dtype=torch.bfloat16)
for _ in range(3): gemm_put(a, b, c, 0, out_dtype=torch.bfloat16, cluster=cluster)
torch.cuda.synchronize()
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_put_2sm -s 1 -c 1 -f -o $OUT/prof_gemm_2sm python /tmp/gemm_one.py 3 > $OUT/ncu_gemm_2sm.log 2>&1; tail -2 $OUT/ncu_gemm_2sm.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_put_kernel -s 1 -c 1 -f -o $OUT/prof_gemm_1sm python /tmp/gemm_one.py 0 > $OUT/ncu_gemm_1sm.log 2>&1; tail -2 $OUT/ncu_gemm_1sm.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_bench_kernel -s 2 -c 1 -f -o $OUT/prof_fused_bench bin/concurency fused --repetitions 4 --commands A D2D > $OUT/ncu_fused_bench.log 2>&1; tail -2 $OUT/ncu_fused_bench.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 2> $OUT/bench_n1.err | grep '^{' > $OUT/bench_n1.json; cut -c1-300 $OUT/bench_n1.json
echo "== r2 call5 done"
