#!/usr/bin/env bash
# Multi-GPU call for the fused tensor-parallel layers (run with gpurun --gpus 2, then --gpus 8; ~5 min of box time
# each, i.e. ~10 and ~40 GPU-minutes):
# exactness against cuBLAS + NCCL, then fused vs stock timings for a small and a Llama-70B-like shape.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/../..}"
OUT=gpurun_out; mkdir -p $OUT
N=${1:-$(nvidia-smi -L | wc -l)}
export HPCP_EXPERIMENTAL=1
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 "$@"; }
# two-slot K-ring (VA/VB double buffer + acks): exactness at P = 2 and 4, then timing next to the default at P = N
timeout 400 python -m pytest tests/test_gpu_multi.py -k "two_slots or pull_ring" -q --timeout 120 2>&1 | tail -4 | tee $OUT/ring_variants_pytest_n$N.txt
for variant in "" "--slots 2" "--pull" "--pull --slots 2"; do   # default (validated) first, then the three new variants
  timeout 120 bin/allreduce -n $N -p 25 --iters 5 $variant --json $OUT/ring_variants_n$N.jsonl | tail -1 | sed "s/^/[$variant] /"
done
run scripts/tp_bench.py --check --mlp --tokens 2048 --out-features 2048 --in-features 2048 --steps 3 2>&1 | tail -3 | tee $OUT/tp_check_n$N.json
# GPU-minutes are charged per GPU: on more than two GPUs only the two headline shapes and one gather granularity
if [ "$N" -le 2 ]; then SHAPES=("8192 8192 8192" "8192 8192 28672" "16384 8192 8192"); CHUNKS="1024 2048"
else SHAPES=("8192 8192 8192" "8192 8192 28672"); CHUNKS="2048"; fi
for shape in "${SHAPES[@]}"; do
  set -- $shape
  run scripts/tp_bench.py --check --mlp --tokens $1 --out-features $2 --in-features $3 2>/dev/null | grep '^{' | tee -a $OUT/tp_bench_n$N.jsonl
  run scripts/tp_bench.py --check --rs-epilogue tma --tokens $1 --out-features $2 --in-features $3 2>/dev/null | grep '^{' | tee -a $OUT/tp_bench_n$N.jsonl
  for chunk in $CHUNKS; do
    run scripts/tp_bench.py --tokens $1 --out-features $2 --in-features $3 --chunk $chunk 2>/dev/null | grep '^{' | sed "s/^{/{\"chunk\": $chunk, /" | tee -a $OUT/tp_bench_n$N.jsonl
  done
done
