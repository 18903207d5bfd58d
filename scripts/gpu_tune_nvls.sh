#!/usr/bin/env bash
# NVLS allreduce tuning sweep (needs >= 2 GPUs with NVSwitch multicast; meant for `gpurun --gpus 8`):
# in-flight reductions per thread x block size x CTAs per SM, at the miniapp's size and at 1 GiB.
# Default configuration (4, 512, 1) measured 0.326 ms at 2^25 floats on 8xB200 = 0.60 of the NVLink roofline.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/nvls_tune.jsonl $OUT/nvls_tune.txt
N=${1:-$(nvidia-smi -L | grep -c '^GPU ')}
for p in 25 28; do
  for u in 1 2 4 8; do
    for t in 256 512 1024; do
      for k in 1 2 4; do
        [ "$t" = 1024 ] && [ "$u" = 8 ] && continue
        [ "$t" = 1024 ] && [ "$k" = 4 ] && continue          # 2 x 1024 threads is the SM limit
        line=$(HPCP_NVLS_UNROLL=$u HPCP_NVLS_THREADS=$t HPCP_NVLS_CTAS_PER_SM=$k timeout 120 \
               ./bin/allreduce -n "$N" -p $p -a --coll nvls --iters 10 --json $OUT/nvls_tune.jsonl 2>&1 | grep Elapsed)
        echo "p=$p unroll=$u threads=$t ctas_per_sm=$k | $line" | tee -a $OUT/nvls_tune.txt
      done
    done
  done
done
sort -t'|' -k2 $OUT/nvls_tune.txt | head -5
