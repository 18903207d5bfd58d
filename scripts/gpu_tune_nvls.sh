#!/usr/bin/env bash
# NVLS allreduce tuning sweep (needs NVSwitch multicast; meant for `gpurun --gpus 8`, ~2-3 min of box time = ~20
# GPU-minutes): at the miniapp's size (2^25 floats) and at 1 GiB (2^28), where NCCL is still ahead.
# Default configuration (unroll 4, 512 threads, one CTA per SM) measured 0.326 ms at 2^25 floats on 8xB200 = 0.60
# of the NVLink roofline; earlier sweeps (148..592 CTAs, access patterns) moved it by <= 10 %.
# New axis: FEWER CTAs than SMs — NCCL drives NVLS from 16-32 CTAs; if the in-switch reduction saturates at modest
# request parallelism, a small grid with deep per-thread unrolling is the better shape.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/nvls_tune.jsonl $OUT/nvls_tune.txt
N=${1:-$(nvidia-smi -L | grep -c '^GPU ')}
run() {  # run <label> <env...> -- <extra args...>
  local label=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  local line
  line=$(env "${envs[@]}" timeout 120 ./bin/allreduce -n "$N" -p $p -a --coll nvls --iters 10 "$@" \
         --json $OUT/nvls_tune.jsonl 2>&1 | grep "Elapsed")
  echo "p=$p $label | $line" | tee -a $OUT/nvls_tune.txt
}
for p in 25 28; do
  for u in 4 8; do
    for k in 1 2; do run "unroll=$u threads=512 ctas_per_sm=$k" HPCP_NVLS_UNROLL=$u HPCP_NVLS_CTAS_PER_SM=$k --; done
    for c in 16 32 64 96; do run "unroll=$u threads=512 ctas=$c" HPCP_NVLS_UNROLL=$u -- --ctas $c; done
  done
  timeout 120 ./bin/allreduce -n "$N" -p $p -a --coll twoshot --iters 10 2>&1 | grep Elapsed | sed "s/^/p=$p twoshot | /" | tee -a $OUT/nvls_tune.txt
done
for p in 25 28; do
  echo "== fastest at p=$p"
  grep "^p=$p " $OUT/nvls_tune.txt | awk -F'\\): ' '{split($2, a, " "); print a[1], "|", $1}' | sort -n | head -4
done
