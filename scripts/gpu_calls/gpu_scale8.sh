#!/usr/bin/env bash
# 8-GPU confirmation run: topology, scaling of the flagship, peer2pear on 4 pairs, allreduce at P=8.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/../..}"
OUT=gpurun_out; mkdir -p $OUT
NGPU=$(nvidia-smi -L | wc -l)
t() { local secs=$1; shift; timeout "$secs" "$@"; }
echo "== $NGPU GPUs" | tee $OUT/scale_summary.txt
t 60 ./bin/topology --matrix 2>&1 | tee $OUT/topology8.txt | head -12 | tee -a $OUT/scale_summary.txt
t 60 ./bin/topology --json > $OUT/topology8.json 2>&1
nvidia-smi topo -m > $OUT/nvidia_smi_topo8.txt 2>&1
for r in 0 3 7; do echo "rank $r -> compact $(./bin/topology --policy compact --rank $r) spread $(./bin/topology --policy spread --rank $r) compact_plan $(./bin/topology --policy compact_plan --rank $r)"; done | tee -a $OUT/scale_summary.txt

# flagship scaling 1,2,4,8 (default engine) + ldst engine at 8
t 200 python bench.py --gpus 1 > $OUT/scale_n1.json 2> $OUT/scale_n1.err; grep '^{' $OUT/scale_n1.json | cut -c1-400 | tee -a $OUT/scale_summary.txt
for n in 2 4 8; do
  [ "$NGPU" -ge $n ] || continue
  t 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2960$n \
      bench.py --gpus $n > $OUT/scale_n$n.json 2> $OUT/scale_n$n.err
  grep '^{' $OUT/scale_n$n.json | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print({k:d[k] for k in ['n_gpus','value','ms_per_step','per_gpu_GBps','overlap_pct','speedup_vs_stock_memcpy','speedup_vs_stock_nccl','e2e','clocks']})" | tee -a $OUT/scale_summary.txt
done
HPCP_BENCH_ENGINE=ldst t 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NGPU --master-addr 127.0.0.1 --master-port 29619 \
    bench.py --gpus $NGPU --no-extras 2>/dev/null | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('ldst engine', {k:d[k] for k in ['n_gpus','value','ms_per_step']})" | tee -a $OUT/scale_summary.txt
HPCP_BENCH_ENGINE=ldst HPCP_BENCH_CTAS=592 t 100 python bench.py --gpus 1 --no-extras 2>/dev/null | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('n1 ldst 592 ctas', d['ms_per_step'])" | tee -a $OUT/scale_summary.txt

# peer2pear on all pairs
rm -f $OUT/p2p8.jsonl
for tr in put get sendrecv memcpy; do
  t 120 ./bin/peer2pear "$tr n$NGPU" -n $NGPU --transport $tr --json $OUT/p2p8.jsonl 2>&1 | tee -a $OUT/scale_summary.txt
done
for mp in spread compact_plan; do
  t 120 ./bin/peer2pear "put $mp" -n $NGPU --mapping $mp --json $OUT/p2p8.jsonl 2>&1 | tee -a $OUT/scale_summary.txt
done
t 120 ./bin/peer2pear "fused-triad n$NGPU" -n $NGPU --fused-triad --engine tma --json $OUT/p2p8.jsonl 2>&1 | tee -a $OUT/scale_summary.txt

# allreduce miniapp at P = NGPU (2^25 floats = 128 MiB per rank)
rm -f $OUT/allreduce8.jsonl
for args in "" "--chunk 131072" "--algo ring-unfused" "-a" "-a --coll twoshot" "--type int" "-a --type int" "-p 28 -a" "-p 28 -a --coll twoshot"; do
  t 180 ./bin/allreduce -n $NGPU --json $OUT/allreduce8.jsonl $args 2>&1 | grep -E "Elapsed|Error|FAILED|NVLS" | tee -a $OUT/scale_summary.txt
done
for algo in ring twoshot nvls nccl ring-nccl; do
  t 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NGPU --master-addr 127.0.0.1 --master-port 29677 \
     -m hpc_patterns_b200.models.allreduce --algo $algo --json $OUT/allreduce8.jsonl 2>&1 | grep -E "Elapsed|rror" | head -3 | tee -a $OUT/scale_summary.txt
done
t 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NGPU --master-addr 127.0.0.1 --master-port 29678 \
     -m hpc_patterns_b200.models.allreduce --algo nccl -p 28 2>&1 | grep -E "Elapsed|rror" | head -3 | tee -a $OUT/scale_summary.txt
echo "== scale done" | tee -a $OUT/scale_summary.txt
