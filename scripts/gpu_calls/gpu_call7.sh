#!/usr/bin/env bash
# 1-GPU call: concurrency tables (reduced env matrix) + fused-mode rechecks after the carveout fix.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/../..}"
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/cuda.log $OUT/cuda.jsonl
LC=(--commands C C --commands C M2D --commands C D2M --commands M2D D2M --commands H2D D2H)
for envs in "HPCP_DEVICE=0" "HPCP_DEVICE=0 CUDA_DEVICE_MAX_CONNECTIONS=1"; do
  ( export $envs; echo "+ export $envs"
    for mode in out_of_order in_order host_threads nowait fused; do
      timeout 120 ./bin/concurency $mode --repetitions 5 "${LC[@]}" --json $OUT/cuda.jsonl
    done ) 2>&1 | tee -a $OUT/cuda.log | grep -E "^##|export"
done
PYTHONPATH=. python -m hpc_patterns_b200.utils.parse $OUT/cuda.log | tee $OUT/cuda_tables.txt
echo "---- tensor-core command and fused-mode rechecks"
timeout 200 ./bin/concurency fused --repetitions 5 --commands T C --commands T H2D --commands T A --commands A H2D --commands C H2D --json $OUT/cuda_t.jsonl 2>&1 | grep -E "^##|Minimum Time|Total Time|Speedup" | tee $OUT/call7_fused.txt
timeout 200 ./bin/concurency out_of_order --repetitions 5 --commands T C --commands T H2D --json $OUT/cuda_t.jsonl 2>&1 | grep -E "^##|Total Time|Speedup" | tee -a $OUT/call7_fused.txt
