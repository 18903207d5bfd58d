#!/usr/bin/env bash
# Round 2, 2-GPU call: the flagship over NVLink (driver-protocol bench, geometry sweep, ncu with NVLink counters),
# K-p2p geometry sweep at 180 MiB, the never-run ring variants, the multi-GPU suite.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/../..}"
OUT=gpurun_out/r2c3; mkdir -p $OUT
N=2
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 "$@"; }
run bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; tail -c 600 $OUT/bench_n2.err; cut -c1-3000 $OUT/bench_n2.json
run bench.py --gpus $N --steps 200 --warmup 5 --no-extras > $OUT/bench_n2_200.json 2>> $OUT/bench_n2.err; cut -c1-500 $OUT/bench_n2_200.json
run scripts/halo_tune.py --out $OUT/halo_tune_n2.jsonl --modes pull push --geometry 16x6 16x7 12x8 8x6 32x6 2>&1 | grep '^{' | cut -c1-260
run scripts/halo_tune.py --out $OUT/halo_tune_n2_rows.jsonl --modes pull push --geometry 16x6 --rows 1 3 5 2>&1 | grep '^{' | cut -c1-260
timeout 900 python -m pytest tests/test_gpu_multi.py -q --timeout 200 -k "not virtual_ranks" > $OUT/pytest_multi_full.txt 2>&1; tail -15 $OUT/pytest_multi_full.txt
for v in "" "--mode push" "--per-step" "--stock memcpy" "--rows 1" "--rows 1 --mode push"; do
  timeout 120 bin/halo -n $N $v --json $OUT/halo_cli_n$N.jsonl 2>&1 | tail -1 | cut -c1-250
done
run scripts/p2p_tune.py --out $OUT/p2p_tune.jsonl --quick 2>&1 | grep -E '^\{|skip' | cut -c1-200
export HPCP_EXPERIMENTAL=1
timeout 600 python -m pytest tests/test_gpu_multi.py -k "two_slots or pull_ring" -q --timeout 200 2>&1 | tail -6 | tee $OUT/pytest_ring_variants.txt
for variant in "" "--slots 2" "--pull" "--pull --slots 2"; do
  timeout 120 bin/allreduce -n $N -p 25 --iters 5 $variant --json $OUT/ring_variants_n$N.jsonl | tail -1 | sed "s/^/[$variant] /"
done
run scripts/tp_bench.py --check --mlp --tokens 2048 --out-features 2048 --in-features 2048 --steps 3 2>&1 | tail -3 | tee $OUT/tp_check_n$N.txt | cut -c1-600
unset HPCP_EXPERIMENTAL
# ncu: one process drives both GPUs; per-step launches never wait for a later launch, so serialisation cannot deadlock
for mode in pull push; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:halo_stencil -s 4 -c 2 -f -o $OUT/prof_halo_${mode}_n2 \
    python scripts/ncu_halo.py --world 2 --devices 0 1 --mode $mode --bytes 50331648 --steps 4 > $OUT/ncu_halo_$mode.log 2>&1; tail -2 $OUT/ncu_halo_$mode.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:copy_ -s 4 -c 4 -f -o $OUT/prof_p2p_n2 \
  python scripts/ncu_p2p.py --bytes 100663296 --reps 2 > $OUT/ncu_p2p.log 2>&1; tail -2 $OUT/ncu_p2p.log
# diagnosis of the >= 3 thread-ranks-per-GPU timeouts (GPU 0 only)
export CUDA_VISIBLE_DEVICES=0
HPCP_TRACE=1 timeout 100 bin/halo -n 3 --bytes 4198400 --rows 3 --steps 3 --iters 1 --stock memcpy > $OUT/diag_halo_stock_n3.txt 2>&1; tail -25 $OUT/diag_halo_stock_n3.txt
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 100 bin/halo -n 3 --bytes 4198400 --rows 3 --steps 3 --iters 1 --stock memcpy 2>&1 | tail -3 | sed 's/^/[maxconn32] /'
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 100 bin/allreduce -n 4 -p 18 --iters 2 2>&1 | tail -3 | sed 's/^/[maxconn32 allreduce n4] /'
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 100 bin/peer2pear v -n 2 --transport memcpy --bytes 4194304 --iters 2 2>&1 | tail -3 | sed 's/^/[maxconn32 p2p memcpy] /'
echo "== r2 call3 done"
