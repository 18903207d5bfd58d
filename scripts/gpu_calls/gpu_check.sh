#!/usr/bin/env bash
# One gpurun call = everything we want to learn from the box (run with --gpus 2 or 8).
# Every step is individually bounded by `timeout`; outputs go to gpurun_out/.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/../..}"
OUT=gpurun_out; mkdir -p $OUT
NGPU=$(nvidia-smi -L | wc -l)
echo "== $NGPU GPUs" | tee $OUT/summary.txt
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee -a $OUT/summary.txt
nvidia-smi topo -m > $OUT/nvidia_smi_topo.txt 2>&1
t() { local secs=$1; shift; timeout "$secs" "$@"; local rc=$?; echo "[rc=$rc] $*" >> $OUT/summary.txt; return $rc; }

t 60 ./bin/topology --matrix > $OUT/topology.txt 2>&1
t 60 ./bin/topology --json > $OUT/topology.json 2>&1
t 60 ./bin/interop_torchless > $OUT/interop.txt 2>&1
t 60 ./bin/interop_driver >> $OUT/interop.txt 2>&1
cat $OUT/interop.txt >> $OUT/summary.txt
if [ "${SKIP_TESTS:-0}" != 1 ]; then
  t 1500 python -m pytest tests -m gpu -q --timeout 300 -x > $OUT/pytest.log 2>&1
  tail -15 $OUT/pytest.log | tee -a $OUT/summary.txt
fi
t 300 python bench.py --gpus 1 --steps 50 --warmup 5 > $OUT/bench_n1.json 2> $OUT/bench_n1.err
cat $OUT/bench_n1.json | tee -a $OUT/summary.txt
for n in 2 4 8; do
  if [ "$NGPU" -ge $n ]; then
    t 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n \
        bench.py --gpus $n --steps 50 --warmup 5 > $OUT/bench_n$n.json 2> $OUT/bench_n$n.err
    grep '^{' $OUT/bench_n$n.json | tee -a $OUT/summary.txt
  fi
done
if [ "$NGPU" -ge 2 ]; then
  rm -f $OUT/p2p.jsonl
  for tr in "put ldst" "put tma" "get ldst" "get tma" "sendrecv ldst" "memcpy ldst"; do
    set -- $tr
    t 300 ./bin/peer2pear "$1-$2" -n 2 --transport $1 --engine $2 --sweep --iters 10 --json $OUT/p2p.jsonl >> $OUT/p2p_sweep.txt 2>&1
  done
  t 120 ./bin/peer2pear "fused-ldst" -n 2 --fused-triad --engine ldst --json $OUT/p2p.jsonl >> $OUT/p2p_sweep.txt 2>&1
  t 120 ./bin/peer2pear "fused-tma" -n 2 --fused-triad --engine tma --json $OUT/p2p.jsonl >> $OUT/p2p_sweep.txt 2>&1
  grep -E "188743680|fused" $OUT/p2p_sweep.txt | tee -a $OUT/summary.txt
  NR=$NGPU
  for args in "" "-a" "-a --coll twoshot" "--algo ring-unfused" "--type int" "-a --type int"; do
    t 180 ./bin/allreduce -n $NR --json $OUT/allreduce.jsonl $args >> $OUT/allreduce.txt 2>&1
  done
  grep Elapsed $OUT/allreduce.txt | tee -a $OUT/summary.txt
fi
t 300 ./bin/concurency fused --json $OUT/concurency.jsonl --commands C C --commands C M2D --commands C D2M --commands M2D D2M --commands H2D D2H --commands C H2D --commands A H2D > $OUT/concurency_fused.txt 2>&1
for mode in in_order out_of_order host_threads nowait; do
  t 300 ./bin/concurency $mode --json $OUT/concurency.jsonl --commands C C --commands C M2D --commands C D2M --commands M2D D2M --commands H2D D2H > $OUT/concurency_$mode.txt 2>&1
done
grep -h "^##" $OUT/concurency_*.txt | tee -a $OUT/summary.txt
echo "== done" | tee -a $OUT/summary.txt
