#!/usr/bin/env bash
# Kernel tuning sweep on 2 GPUs + ncu captures on 1 GPU.  Outputs -> gpurun_out/.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/../..}"
OUT=gpurun_out; mkdir -p $OUT
B=188743680
t() { local secs=$1; shift; timeout "$secs" "$@"; }

# 1. put copy kernel: vector width x layout x grid x block
rm -f $OUT/tune_p2p.jsonl
run() { local label=$1; shift; t 60 ./bin/peer2pear "$label" -n 2 --bytes $B --iters 5 --no-verify --json $OUT/tune_p2p.jsonl "$@" > /dev/null 2>> $OUT/tune.err; }
for vec in 16 32; do for blocked in 0 1; do for ctas in 32 74 148 296; do for threads in 512 1024; do
  extra=""; [ $blocked = 1 ] && extra="--blocked"
  run "put v$vec b$blocked c$ctas t$threads u4" --transport put --vec $vec --ctas $ctas --threads $threads --unroll 4 $extra
done; done; done; done
for vec in 16 32; do for ctas in 148 296; do
  run "get v$vec c$ctas u4" --transport get --vec $vec --ctas $ctas --unroll 4
done; done
# 2. TMA engine: stage size x stages x grid
for skb in 16 32 64; do for st in 3 6; do for ctas in 148 296; do
  [ $((skb*st)) -gt 200 ] && continue
  run "put tma k$skb s$st c$ctas" --transport put --engine tma --stage-kb $skb --stages $st --ctas $ctas
done; done; done
# 3. fused triad + put
for vec in 16 32; do for blocked in 0 1; do for ctas in 148 296 444; do
  extra=""; [ $blocked = 1 ] && extra="--blocked"
  run "fused v$vec b$blocked c$ctas u2" --fused-triad --vec $vec --ctas $ctas --unroll 2 $extra
done; done; done
for ctas in 148 296; do run "fused v16 b0 c$ctas u4" --fused-triad --vec 16 --ctas $ctas --unroll 4; run "fused tma c$ctas" --fused-triad --engine tma --ctas $ctas; done
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/tune_p2p.jsonl')]
rows.sort(key=lambda r:-r['bi_GBps'])
print("top by bidirectional GB/s")
for r in rows[:25]: print("%-34s uni %.1f bi %.1f"%(r['label'],r['uni_GBps'],r['bi_GBps']))
rows.sort(key=lambda r:-r['uni_GBps'])
print("top by unidirectional GB/s")
for r in rows[:15]: print("%-34s uni %.1f bi %.1f"%(r['label'],r['uni_GBps'],r['bi_GBps']))
PY
# 4. allreduce, native + python (incl. the NCCL baselines)
for args in "" "--chunk 8192" "--chunk 131072" "-a" "-a --coll twoshot" "--algo ring-unfused"; do
  t 120 ./bin/allreduce -n 2 --json $OUT/allreduce2.jsonl $args 2>&1 | grep -E "Elapsed|Error|FAILED"
done
for algo in ring twoshot nvls nccl ring-nccl; do
  t 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 \
     -m hpc_patterns_b200.models.allreduce --algo $algo --json $OUT/allreduce2.jsonl 2>&1 | grep -E "Elapsed|Error|rror:" | head -3
done
# 5. concurrency fused re-check
t 200 ./bin/concurency fused --commands H2D D2H --commands A H2D --commands C D2P --commands A D2P --commands D2P P2D 2>&1 | grep -E "^##|Minimum|Speedup"
# 6. bench at 1 and 2 GPUs
t 200 python bench.py --gpus 1 | tee $OUT/bench_n1.json
t 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29578 bench.py --gpus 2 | grep '^{' | tee $OUT/bench_n2.json
# 7. ncu: launch list + full capture of the flagship kernel (1 GPU, loop-back)
t 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 5 --warmup 3 --e2e-steps 1 --no-extras > $OUT/ncu_launches.log 2>&1
t 400 ncu --set full --clock-control none --import-source on -k regex:triad_put -s 3 -c 2 -f -o $OUT/prof_triad_put \
    python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-extras > $OUT/ncu_triad.log 2>&1
t 400 ncu --set full --clock-control none --import-source on -k regex:triad_put -s 3 -c 2 -f -o $OUT/prof_triad_put_tma \
    python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-extras --engine tma > $OUT/ncu_triad_tma.log 2>&1
ls -la $OUT/*.ncu-rep
echo "== tune done"
