#!/usr/bin/env bash
# 2-GPU follow-up: hybrid put+get transport, rows=8 bench at N=2 and N=1, the tests that failed in call 3.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/../..}"
OUT=gpurun_out/r2c3b; mkdir -p $OUT
N=2
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 "$@"; }
run scripts/p2p_tune.py --out $OUT/p2p_hybrid.jsonl --hybrid 2>&1 | grep -E '^\{|skip' | cut -c1-220
run bench.py --gpus $N --steps 20 --warmup 5 2> $OUT/bench_n2.err | grep '^{' > $OUT/bench_n2.json; tail -c 300 $OUT/bench_n2.err; cut -c1-600 $OUT/bench_n2.json
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras 2> $OUT/bench_n1.err | grep '^{' > $OUT/bench_n1.json; cut -c1-400 $OUT/bench_n1.json
timeout 600 python -m pytest tests/test_gpu_multi.py -q --timeout 200 -k "torchrun_workers or two_slots or pull_ring or bench_two" > $OUT/pytest_retry.txt 2>&1; tail -8 $OUT/pytest_retry.txt
timeout 300 python -m pytest tests/test_gpu_halo.py -q --timeout 120 2>&1 | tail -4 | tee $OUT/pytest_halo.txt
echo "== r2 call3b done"
