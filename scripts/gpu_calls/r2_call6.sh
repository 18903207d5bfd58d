#!/usr/bin/env bash
# Round 2, last 1-GPU call: the final tree — whole GPU suite, smoke(), driver-protocol bench, ncu of the two GEMM loops.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/../..}"
OUT=gpurun_out/r2c6; mkdir -p $OUT
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > $OUT/pytest_gpu_full.txt 2>&1; tail -6 $OUT/pytest_gpu_full.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.txt
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 2> $OUT/bench_n1.err | grep '^{' > $OUT/bench_n1.json; cut -c1-260 $OUT/bench_n1.json
timeout 100 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_reference_arm.json 2>&1; cut -c1-400 $OUT/bench_reference_arm.json
cat > $OUT/gemm_one.py <<'PY'
import sys, torch
from hpc_patterns_b200.ops.gemm import gemm_put
cluster = int(sys.argv[1])
a = torch.randn(8192, 4096, device='cuda').bfloat16(); b = torch.randn(8192, 4096, device='cuda').bfloat16()
c = torch.empty(8192, 8192, device='cuda', dtype=torch.bfloat16)
for _ in range(3): gemm_put(a, b, c, 0, out_dtype=torch.bfloat16, cluster=cluster)
torch.cuda.synchronize()
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_put_2sm -s 1 -c 1 -f -o $OUT/prof_gemm_2sm python $OUT/gemm_one.py 3 > $OUT/ncu_gemm_2sm.log 2>&1; tail -1 $OUT/ncu_gemm_2sm.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_put_kernel -s 1 -c 1 -f -o $OUT/prof_gemm_1sm python $OUT/gemm_one.py 0 > $OUT/ncu_gemm_1sm.log 2>&1; tail -1 $OUT/ncu_gemm_1sm.log
echo "== r2 call6 done"
