#!/usr/bin/env bash
# Round 2, the ONE 8-GPU call (charged 8x): flagship at N=8 under the driver's protocol, native halo CLI, NVLS / two-shot
# sweep against NCCL at 128 MiB and 1 GiB, ring variants, peer2pear on 4 pairs, typed reductions, fused TP check.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/../..}"
OUT=gpurun_out/r2c4; mkdir -p $OUT
N=8
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 "$@"; }
run bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n8.json 2> $OUT/bench_n8.err; tail -c 400 $OUT/bench_n8.err; cut -c1-2800 $OUT/bench_n8.json
for v in "" "--mode push" "--rows 1" "--rows 1 --mode push" "--stock memcpy"; do
  timeout 90 bin/halo -n $N $v --json $OUT/halo_cli_n8.jsonl 2>&1 | tail -1 | cut -c1-260
done
# one-launch collectives: fewer CTAs than SMs and deeper unrolling (NCCL drives NVLS from 16-32 CTAs)
for p in 25 28; do
  for u in 4 8; do
    for c in 32 64 148 296; do
      line=$(HPCP_NVLS_UNROLL=$u timeout 90 bin/allreduce -n $N -p $p -a --coll nvls --iters 10 --ctas $c --json $OUT/nvls_tune.jsonl 2>&1 | grep Elapsed)
      echo "p=$p nvls unroll=$u ctas=$c | $line" | cut -c1-200 | tee -a $OUT/nvls_tune.txt
    done
  done
  for c in 148 296 592; do
    line=$(timeout 90 bin/allreduce -n $N -p $p -a --coll twoshot --iters 10 --ctas $c --json $OUT/nvls_tune.jsonl 2>&1 | grep Elapsed)
    echo "p=$p twoshot ctas=$c | $line" | cut -c1-200 | tee -a $OUT/nvls_tune.txt
  done
  run -m hpc_patterns_b200 allreduce --algo nccl -p $p --iters 10 2>&1 | grep Elapsed | sed "s/^/p=$p NCCL | /" | cut -c1-200 | tee -a $OUT/nvls_tune.txt
done
for variant in "" "--slots 2" "--pull" "-a --type int" "--type double" "--type long" "-a --type double"; do
  timeout 90 bin/allreduce -n $N -p 25 --iters 5 $variant --json $OUT/allreduce_n8.jsonl 2>&1 | tail -1 | sed "s/^/[$variant] /" | cut -c1-220
done
for t in put get memcpy; do
  timeout 90 bin/peer2pear "p2p8 $t" -n $N --transport $t --engine tma --json $OUT/p2p_n8.jsonl 2>&1 | tail -2 | cut -c1-200
done
HPCP_EXPERIMENTAL=1 run scripts/tp_bench.py --check --mlp --tokens 8192 --out-features 8192 --in-features 8192 2>&1 | grep '^{' | tee $OUT/tp_n8.json | cut -c1-900
echo "== r2 call4 done"
