#!/usr/bin/env bash
# Multi-GPU call for the fused tensor-parallel layers (run with gpurun --gpus 2, then --gpus 8):
# exactness against cuBLAS + NCCL, then fused vs stock timings for a small and a Llama-70B-like shape.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT
N=${1:-$(nvidia-smi -L | wc -l)}
export HPCP_EXPERIMENTAL=1
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 "$@"; }
run scripts/tp_bench.py --check --m 2048 --n 2048 --k 2048 --steps 3 2>&1 | tail -3 | tee $OUT/tp_check_n$N.json
for shape in "8192 8192 8192" "8192 8192 28672" "16384 8192 8192"; do
  set -- $shape
  run scripts/tp_bench.py --check --m $1 --n $2 --k $3 2>/dev/null | grep '^{' | tee -a $OUT/tp_bench_n$N.jsonl
  for chunk in 1024 2048; do
    run scripts/tp_bench.py --m $1 --n $2 --k $3 --chunk $chunk 2>/dev/null | grep '^{' | sed "s/^{/{\"chunk\": $chunk, /" | tee -a $OUT/tp_bench_n$N.jsonl
  done
done
